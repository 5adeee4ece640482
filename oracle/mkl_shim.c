/* Test infrastructure only (oracle/): the non-DFTI pieces of MKL the reference C
 * core calls - aligned allocation and the two BLAS-1 conjugated dot products
 * (call sites: reference projector.c:269,1016,1025; pseudoprojector.c:86; density.c:223).
 * Sequential loops, accumulating in the precision of the routine (zdotc: double,
 * cdotc: float) exactly as the BLAS interface prescribes. */
#include <complex.h>
#include <stdlib.h>
#include <string.h>

void *mkl_malloc(size_t size, int align) {
  void *p = NULL;
  if (align < (int)sizeof(void *)) align = sizeof(void *);
  if (posix_memalign(&p, (size_t)align, size ? size : 1)) return NULL;
  return p;
}
void *mkl_calloc(size_t n, size_t size, int align) {
  void *p = mkl_malloc(n * size, align);
  if (p) memset(p, 0, n * size);
  return p;
}
void mkl_free(void *p) { free(p); }
void mkl_free_buffers(void) {}

void cblas_zdotc_sub(const int n, const void *x, const int incx, const void *y,
                     const int incy, void *dotc) {
  const double complex *a = (const double complex *)x, *b = (const double complex *)y;
  double complex s = 0;
  for (int i = 0; i < n; i++) s += conj(a[(long)i * incx]) * b[(long)i * incy];
  *(double complex *)dotc = s;
}
void cblas_cdotc_sub(const int n, const void *x, const int incx, const void *y,
                     const int incy, void *dotc) {
  const float complex *a = (const float complex *)x, *b = (const float complex *)y;
  float complex s = 0;
  for (int i = 0; i < n; i++) s += conjf(a[(long)i * incx]) * b[(long)i * incy];
  *(float complex *)dotc = s;
}
