/* cuhd_util.h -- include-path compatibility with gpuhd/include/cuhd_util.h: everything lives in cuhd.h */
#include "cuhd.h"
