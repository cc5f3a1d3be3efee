// llpf_rtc_compat.h — lets the engine headers compile under NVRTC as well as nvcc.
// NVRTC has no host C library headers; under nvcc this file is just the three usual includes.
// (Probe: scripts/nvrtc_probe.py compiles k_engine<4,2,0,0> to an sm_100a cubin at run time in ~10 s — the basis of the
//  planned user-function hook)
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>
#else
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#ifndef DBL_MAX
#define DBL_MAX 1.7976931348623157e+308
#endif
#endif
