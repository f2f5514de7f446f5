/* cuhd_gpu_decoder.h -- include-path compatibility with gpuhd/include/cuhd_gpu_decoder.h: everything lives in cuhd.h */
#include "cuhd.h"
