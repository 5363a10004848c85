// Version / architecture queries of the C-ABI (include/dcl_b200.h).
#include "common.cuh"
#include "../../include/dcl_b200.h"

DCL_API int dcl_b200_abi_version(void) { return DCL_B200_ABI_VERSION; }
DCL_API int dcl_b200_arch(void) { return 100; }
