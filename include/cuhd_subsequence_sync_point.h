/* cuhd_subsequence_sync_point.h -- include-path compatibility with gpuhd/include/cuhd_subsequence_sync_point.h: everything lives in cuhd.h */
#include "cuhd.h"
