/* cuhd_definitions.h -- include-path compatibility with gpuhd/include/cuhd_definitions.h: everything lives in cuhd.h */
#include "cuhd.h"
