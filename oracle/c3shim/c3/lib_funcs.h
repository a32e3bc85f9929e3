/* C3 stand-in: only the function_class enum (util.c:113-147). */
#ifndef C3SHIM_FUNCS_H
#define C3SHIM_FUNCS_H
enum function_class { CONSTANT, PIECEWISE, POLYNOMIAL, LINELM, CONSTELM, KERNEL };
#endif
