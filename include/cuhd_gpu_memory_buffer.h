/* cuhd_gpu_memory_buffer.h -- include-path compatibility with gpuhd/include/cuhd_gpu_memory_buffer.h: everything lives in cuhd.h */
#include "cuhd.h"
