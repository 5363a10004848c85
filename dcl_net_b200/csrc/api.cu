// Version / architecture queries of the C-ABI (include/dcl_b200.h).
#include "common.cuh"
#include "../../include/dcl_b200.h"

DCL_API int dcl_b200_abi_version(void) { return DCL_B200_ABI_VERSION; }
DCL_API int dcl_b200_arch(void) { return 100; }

unsigned long long g_dcl_kernel_launches = 0;
DCL_API unsigned long long dcl_b200_launch_count(void) { return g_dcl_kernel_launches; }
