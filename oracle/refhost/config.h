/* Hand-written build configuration for compiling the UNMODIFIED coin-or/Ipopt
 * host library from /root/reference with oracle/refhost/Makefile (we do not run
 * the reference's autotools).  Mirrors the macros of src/Common/config.h.in that
 * the sources actually test (grep'd: IpUtils.cpp, IpJournalist.cpp,
 * IpLibraryLoader.cpp, IpIpoptAlg.cpp, IpBlas.cpp, IpLapack.cpp).
 * TEST/BASELINE INFRASTRUCTURE - not product source. */
#ifndef B200_REFHOST_CONFIG_H
#define B200_REFHOST_CONFIG_H
#define HAVE_CMATH 1
#define HAVE_CFLOAT 1
#define HAVE_MATH_H 1
#define HAVE_FLOAT_H 1
#define HAVE_STDIO_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRING_H 1
#define HAVE_STDINT_H 1
#define HAVE_INTTYPES_H 1
#define HAVE_UNISTD_H 1
#define HAVE_DLFCN_H 1
#define HAVE_SYS_STAT_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_VSNPRINTF 1
#define STDC_HEADERS 1
#define IPOPT_HAS_VA_COPY 1
#define IPOPT_HAS_DRAND48 1
#define IPOPT_HAS_RAND 1
#define IPOPT_HAS_STD__RAND 1
#define IPOPT_C_FINITE std::isfinite
#define IPOPT_HAS_LAPACK 1
#define IPOPT_HAS_LINEARSOLVERLOADER 1
#define IPOPT_CHECKLEVEL 0
#define IPOPT_VERBOSITY 0
#define IPOPT_VERSION "3.14.15"
#define IPOPT_VERSION_MAJOR 3
#define IPOPT_VERSION_MINOR 14
#define IPOPT_VERSION_RELEASE 15
#define F77_FUNC(name,NAME) name ## _
#define F77_FUNC_(name,NAME) name ## _
#define IPOPT_LAPACK_FUNC(name,NAME) name ## _
#define IPOPT_LAPACK_FUNC_(name,NAME) name ## _
#define IPOPT_HSL_FUNC(name,NAME) name ## _
#define IPOPT_HSL_FUNC_(name,NAME) name ## _
#define IPOPTLIB_EXPORT __attribute__((__visibility__("default")))
#define SIPOPTLIB_EXPORT
#define IPOPTAMPLINTERFACELIB_EXPORT
#define HSLLIB_EXPORT
#define SIZEOF_INT_P 8
#ifndef IPOPT_FORTRAN_INTEGER_TYPE
#define IPOPT_FORTRAN_INTEGER_TYPE ipindex
#endif
#endif
