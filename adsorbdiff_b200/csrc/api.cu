// Library-level entry points of the C ABI.
#include "common.cuh"

extern "C" int adk_abi_version(void) { return ADK_ABI_VERSION; }

extern "C" int adk_init(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    int rc;
    if ((rc = adk_neighbors_set_attrs()) != 0) return rc;
    if ((rc = adk_message_set_attrs()) != 0) return rc;
    if ((rc = adk_linear_set_attrs()) != 0) return rc;
    if ((rc = adk_linear_tc_set_attrs()) != 0) return rc;
    if ((rc = adk_message_mma_set_attrs()) != 0) return rc;
    if ((rc = adk_message_t5_set_attrs()) != 0) return rc;
    if ((rc = adk_message_bwd_set_attrs()) != 0) return rc;
    return 0;
}
