/* cuhd_gpu_codetable.h -- include-path compatibility with gpuhd/include/cuhd_gpu_codetable.h: everything lives in cuhd.h */
#include "cuhd.h"
