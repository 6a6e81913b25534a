// abfit_rtc_prelude.h — what the device headers need from <stdint.h> / <cuda_runtime.h> when NVRTC compiles them
// (NVRTC provides the CUDA built-ins and size_t, but no host headers).
#pragma once
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
