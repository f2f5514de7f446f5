/* cuhd_constants.h -- include-path compatibility with gpuhd/include/cuhd_constants.h: everything lives in cuhd.h */
#include "cuhd.h"
