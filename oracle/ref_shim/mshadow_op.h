/* parse-only stand-in, see shim_common.h */
#include "shim_common.h"
