/* Public-side configuration header (what `make install` would have written as
 * IpoptConfig.h); used when compiling plugins / drivers against oracle/_ref. */
#ifndef __CONFIG_IPOPT_H__
#define __CONFIG_IPOPT_H__
#define IPOPT_VERSION "3.14.15"
#define IPOPT_VERSION_MAJOR 3
#define IPOPT_VERSION_MINOR 14
#define IPOPT_VERSION_RELEASE 15
#define IPOPT_CHECKLEVEL 0
#define IPOPT_VERBOSITY 0
#ifndef IPOPT_FORTRAN_INTEGER_TYPE
#define IPOPT_FORTRAN_INTEGER_TYPE ipindex
#endif
#ifndef IPOPTLIB_EXPORT
#define IPOPTLIB_EXPORT
#endif
#ifndef SIPOPTLIB_EXPORT
#define SIPOPTLIB_EXPORT
#endif
#endif
