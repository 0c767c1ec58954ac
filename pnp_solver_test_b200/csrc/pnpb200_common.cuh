// pnpb200_common.cuh -- error plumbing and host helpers shared by the translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/pnpb200.h"

namespace pnpb200 {

void set_last_error(const char* where, cudaError_t e);   // defined in pnpb200_solve.cu

#define PNP_CUDA_OK(call)                                                   \
    do {                                                                    \
        cudaError_t _e = (call);                                            \
        if (_e != cudaSuccess) {                                            \
            ::pnpb200::set_last_error(#call, _e);                           \
            return (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver) ? PNPB200_ENODEVICE \
                                                                                  : PNPB200_ECUDA;   \
        }                                                                   \
    } while (0)

// np.linalg.inv(K) for the 3x3 camera matrix (PNP_SOLVER_LIB.py:2596, :2811), on the host.
inline void host_inv3(const double* M, double* out)
{
    const double d = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) +
                     M[2] * (M[3] * M[7] - M[4] * M[6]);
    const double id = 1.0 / d;
    out[0] = (M[4] * M[8] - M[5] * M[7]) * id; out[1] = (M[2] * M[7] - M[1] * M[8]) * id; out[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    out[3] = (M[5] * M[6] - M[3] * M[8]) * id; out[4] = (M[0] * M[8] - M[2] * M[6]) * id; out[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    out[6] = (M[3] * M[7] - M[4] * M[6]) * id; out[7] = (M[1] * M[6] - M[0] * M[7]) * id; out[8] = (M[0] * M[4] - M[1] * M[3]) * id;
}

struct DeviceProps {
    int device, sm_count, cc_major, cc_minor, max_smem_optin;
    size_t total_mem;
};
int get_device_props(DeviceProps* out);   // cached per device; defined in pnpb200_solve.cu

}  // namespace pnpb200
