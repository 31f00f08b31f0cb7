// Configuration.h (shim) -- what /root/reference/src/Configuration.h gives its includers, for code written against the
// reference that is compiled against rapidnet-b200: the two scalar typedefs (:30-31), `using namespace std` (:33) and the
// three check macros (:38-81: print where, then exit).  Same names and behaviour, own wording.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>

#include <cuda_runtime.h>

#include "rapidnet_host.hpp"

using namespace std;
using rapidnet::real_t;
using rapidnet::uint_t;

#define _CUDA(call)                                                                                              \
    do {                                                                                                         \
        const cudaError_t rn_err_ = (call);                                                                      \
        if (rn_err_ != cudaSuccess) {                                                                            \
            std::cerr << "CUDA Error: \nFile = " << __FILE__ << "\nLine = " << __LINE__ << " \nReason = " << cudaGetErrorString(rn_err_); \
            cudaDeviceReset();                                                                                   \
            std::exit(EXIT_FAILURE);                                                                             \
        }                                                                                                        \
    } while (0)

#define _CUBLAS(call)                                                                                            \
    do {                                                                                                         \
        const int rn_st_ = (int)(call);                                                                          \
        if (rn_st_ != 0) {                                                                                       \
            std::cerr << "CUBLAS Error: \nFile = " << __FILE__ << "\nLine = " << __LINE__ << " \nReason = " << rn_st_; \
            cudaDeviceReset();                                                                                   \
            std::exit(EXIT_FAILURE);                                                                             \
        }                                                                                                        \
    } while (0)

#define _ASSERT(cond)                                                                                            \
    do {                                                                                                         \
        if (!(cond)) {                                                                                           \
            std::cerr << " Error: \nFile = " << __FILE__ << "\nLine = " << __LINE__ << "\n";                    \
            std::exit(EXIT_FAILURE);                                                                             \
        }                                                                                                        \
    } while (0)
