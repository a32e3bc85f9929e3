/* C3 stand-in (test infrastructure only).
 *
 * The reference links against the third-party C3 library
 * (github.com/goroda/Compressed-Continuous-Computation, unpinned, absent
 * from /root/reference and from this image).  This header declares the few
 * array helpers the reference's C sources call so that those files compile
 * UNMODIFIED, in place, into oracle/_ref/.  Semantics restated from C3's
 * public array.h; none of this is product code.
 */
#ifndef C3SHIM_ARRAY_H
#define C3SHIM_ARRAY_H
#include <stddef.h>
#include <stdio.h>

struct c3Vector { size_t size; double *elem; };

double  *calloc_double(size_t n);
size_t  *calloc_size_t(size_t n);
int     *calloc_int(size_t n);
double **malloc_dd(size_t n);
void     free_dd(size_t n, double **a);
double  *linspace(double lb, double ub, size_t n);
double   randu(void);
void     dprint(size_t n, const double *a);
void     iprint(size_t n, const int *a);
void     iprint_sz(size_t n, const size_t *a);
#endif
