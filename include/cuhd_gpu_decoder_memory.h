/* cuhd_gpu_decoder_memory.h -- include-path compatibility with gpuhd/include/cuhd_gpu_decoder_memory.h: everything lives in cuhd.h */
#include "cuhd.h"
