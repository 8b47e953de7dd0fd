/* Hand-written stand-in for the cmake-generated anakin_config.h (reference
 * cmake/config/anakin_config.h.in): just enough switches to compile saber/core +
 * the reference's naive test oracle (test/saber/conv_func_helper.h) for the x86
 * host target.  Used only to build oracle/_ref/. */
#ifndef _ANAKIN_CONFIGURATION_HEADER_GUARD_H_
#define _ANAKIN_CONFIGURATION_HEADER_GUARD_H_
#define ANAKIN_VERSION "2.0-ref"
#define ANAKIN_TYPE_FP32
#define USE_X86_PLACE
#define USE_OPENMP
#define USE_LOGGER
#define PLATFORM_POSIX
#endif
