/* cuhd_gpu_input_buffer.h -- include-path compatibility with gpuhd/include/cuhd_gpu_input_buffer.h: everything lives in cuhd.h */
#include "cuhd.h"
