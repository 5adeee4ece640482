/* Test infrastructure only (oracle/): a minimal stand-in for Intel MKL's <mkl.h>
 * so that the UNMODIFIED reference C sources under /root/reference/pawpyseed/core
 * compile here. The DFTI entry points are the genuine MKL ones exported by
 * libtorch_cpu.so (MKL is statically embedded there); only the prototypes and
 * the enum values live in this header. BLAS-1 dots and mkl_malloc are provided by
 * mkl_shim.c. Nothing in the product path includes this file. */
#ifndef PAWB200_ORACLE_MKL_SHIM_H
#define PAWB200_ORACLE_MKL_SHIM_H
#include <complex.h>
#include <stddef.h>

#define MKL_LONG long
typedef void *DFTI_DESCRIPTOR_HANDLE;
#ifndef MKL_Complex16
#define MKL_Complex16 double complex
#endif
#ifndef MKL_Complex8
#define MKL_Complex8 float complex
#endif

/* values from MKL's mkl_dfti.h (public ABI constants) */
enum DFTI_CONFIG_PARAM {
  DFTI_FORWARD_DOMAIN = 0, DFTI_DIMENSION = 1, DFTI_LENGTHS = 2, DFTI_PRECISION = 3,
  DFTI_FORWARD_SCALE = 4, DFTI_BACKWARD_SCALE = 5, DFTI_NUMBER_OF_TRANSFORMS = 7,
  DFTI_PLACEMENT = 11
};
enum DFTI_CONFIG_VALUE {
  DFTI_COMPLEX = 32, DFTI_REAL = 33, DFTI_SINGLE = 35, DFTI_DOUBLE = 36,
  DFTI_INPLACE = 43, DFTI_NOT_INPLACE = 44
};

MKL_LONG DftiCreateDescriptor_d_1d(DFTI_DESCRIPTOR_HANDLE *, enum DFTI_CONFIG_VALUE, MKL_LONG);
MKL_LONG DftiCreateDescriptor_d_md(DFTI_DESCRIPTOR_HANDLE *, enum DFTI_CONFIG_VALUE, MKL_LONG, MKL_LONG *);
MKL_LONG DftiSetValue(DFTI_DESCRIPTOR_HANDLE, enum DFTI_CONFIG_PARAM, ...);
MKL_LONG DftiCommitDescriptor(DFTI_DESCRIPTOR_HANDLE);
MKL_LONG DftiComputeForward(DFTI_DESCRIPTOR_HANDLE, void *, ...);
MKL_LONG DftiComputeBackward(DFTI_DESCRIPTOR_HANDLE, void *, ...);
MKL_LONG DftiFreeDescriptor(DFTI_DESCRIPTOR_HANDLE *);
char *DftiErrorMessage(MKL_LONG);

/* the reference only ever asks for DFTI_DOUBLE / DFTI_COMPLEX, 1-D (sbt.c) or 3-D (linalg.c) */
static inline MKL_LONG pawb200_dfti_create(DFTI_DESCRIPTOR_HANDLE *h, int prec, int dom,
                                           MKL_LONG dim, const void *sizes, MKL_LONG size1) {
  (void)prec;
  if (dim == 1) return DftiCreateDescriptor_d_1d(h, (enum DFTI_CONFIG_VALUE)dom, size1);
  return DftiCreateDescriptor_d_md(h, (enum DFTI_CONFIG_VALUE)dom, dim, (MKL_LONG *)sizes);
}
/* dim==1 call sites pass a scalar MKL_LONG, dim==3 ones pass MKL_LONG[3] */
#define DftiCreateDescriptor(h, prec, dom, dim, sz)                                         \
  _Generic((sz), MKL_LONG: pawb200_dfti_create(h, prec, dom, dim, NULL, (MKL_LONG)(size_t)(sz)), \
           default: pawb200_dfti_create(h, prec, dom, dim, (const void *)(size_t)(sz), 0))

void *mkl_malloc(size_t size, int align);
void *mkl_calloc(size_t n, size_t size, int align);
void mkl_free(void *p);
void mkl_free_buffers(void);

void cblas_zdotc_sub(const int n, const void *x, const int incx, const void *y, const int incy, void *dotc);
void cblas_cdotc_sub(const int n, const void *x, const int incx, const void *y, const int incy, void *dotc);
#endif
