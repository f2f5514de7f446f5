/* cuhd_codetable.h -- include-path compatibility with gpuhd/include/cuhd_codetable.h: everything lives in cuhd.h */
#include "cuhd.h"
