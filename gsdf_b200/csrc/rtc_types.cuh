// rtc_types.cuh -- what the headers of the interpreter need from <stdint.h> when they are compiled at run time by NVRTC
// (specialised kernels, jit.cu): NVRTC has no system headers. Device builds by nvcc never include this file.
#pragma once
#ifdef __CUDACC_RTC__
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
// <math.h>'s constants, same bit patterns as glibc's (quiet NaN 0x7fc00000)
#define NAN __int_as_float(0x7fc00000)
#define INFINITY __int_as_float(0x7f800000)
#endif
