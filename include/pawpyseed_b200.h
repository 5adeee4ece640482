/* pawpyseed_b200 - C ABI of the B200-native PAW band-projection engine.
 *
 * Drop-in boundary: every entry point below has the argument list of the reference C
 * function that pawpyseed's Cython shim binds (pawpyseed/core/pawpyc_extern.pxd, generated
 * from the C headers cited per function; call sites in pawpyseed/core/pawpyc.pyx), with
 * the prefix `pawb200_` so both libraries can live in one process.  INTEGRATION.md shows
 * the three-line change to pawpyc_extern.pxd that re-points the shim.
 *
 * Conventions (identical to the reference):
 *   - all arrays are C-contiguous host buffers; `double complex*` results are interleaved
 *     (re,im) complex128;
 *   - site lists may be NULL when their length is 0 (pawpyc.pyx:666-671);
 *   - result index of a projection row is  b*NK + kappa,  kappa = k + s*nwk,  NK = nwk*nspin
 *     (pseudoprojector.h:23-27);
 *   - grids are x-slowest / z-fastest (linalg.h:14-19);
 *   - pawb200_compensation_terms ACCUMULATES (+=) into `overlap` (projector.c:910-959),
 *     pawb200_pseudoprojection overwrites (pseudoprojector.c:87).
 * Differences, all deliberate:
 *   - pswf_t / ppot_t are opaque; wavefunction data lives in GPU memory (HBM);
 *   - nothing calls exit(): failures set a thread-local message readable through
 *     pawb200_last_error() (the reference prints and exit(-1)s, utils.c:1100-1118);
 *   - the pseudo overlap is accumulated in FP64 (the reference uses a single-precision
 *     cblas_cdotc_sub, pseudoprojector.c:86);
 *   - there is NO CPU fallback: every compute entry point fails with an error if no
 *     CUDA device is usable.
 * Threading: like the reference (one caller thread, pswf_t mutated by overlap_setup_*), the
 *   library is not re-entrant - one main CUDA stream, shared scratch buffers, stream-ordered
 *   pools.  Every entry point that touches device state takes one process-wide lock, so
 *   concurrent callers are serialised.  One process drives one GPU (the current device).
 */
#ifndef PAWPYSEED_B200_H
#define PAWPYSEED_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
typedef struct { double re, im; } pawb200_c128;   /* layout of C99 `double complex` */
typedef struct { float re, im; } pawb200_c64;     /* layout of C99 `float complex`  */
#else
#include <complex.h>
typedef double complex pawb200_c128;
typedef float complex pawb200_c64;
#endif

typedef struct pawb200_pswf pawb200_pswf_t;   /* replaces pswf_t  (utils.h:117-141) */
typedef struct pawb200_density_ft pawb200_density_ft_t;   /* replaces density_ft_elem_t* (momentum.h:29-33) */
typedef struct pawb200_ppot pawb200_ppot_t;   /* replaces ppot_t* (utils.h:49-70), a whole element list */

/* ---- status ------------------------------------------------------------------------- */
/* NULL when the last call on this thread succeeded, else a message. */
const char *pawb200_last_error(void);
void pawb200_clear_error(void);
/* 0 when a CUDA device of compute capability 10.x is usable; sets the error otherwise. */
int pawb200_device_check(void);
const char *pawb200_version(void);

/* ---- reader (reader.h:46,52; reader.c:129-315) --------------------------------------- */
pawb200_pswf_t *pawb200_read_wavefunctions(const char *filename, const double *kpt_weights);
pawb200_pswf_t *pawb200_read_wavefunctions_from_str(const char *start, const double *kpt_weights);
void pawb200_free_pswf(pawb200_pswf_t *wf);                                   /* utils.h:242 */

/* ---- k-point desymmetrisation (utils.h:388-393; utils.c:829-1098), SURVEY 8 row f2 --------------- */
/* New wavefunction with num_kpts k-points: k-point q is ops[q] (3x3, reciprocal fractional) applied to
 * source k-point maps[q] (time-reversed when trs[q] == 1) with fractional translation drs[q]; weights kws.
 * The plane-wave remap + phase runs on the GPU; the result lives in HBM like any other wavefunction. */
pawb200_pswf_t *pawb200_expand_symm_wf(pawb200_pswf_t *rwf, int num_kpts, const int *maps,
                                       const double *ops, const double *drs, const double *kws,
                                       const int *trs);

/* ---- accessors (utils.h:257-279) ----------------------------------------------------- */
int pawb200_get_nband(pawb200_pswf_t *wf);
int pawb200_get_nwk(pawb200_pswf_t *wf);
int pawb200_get_nspin(pawb200_pswf_t *wf);
int pawb200_is_ncl(pawb200_pswf_t *wf);
double pawb200_get_encut(pawb200_pswf_t *wf);
double pawb200_get_energy(pawb200_pswf_t *wf, int band, int kpt, int spin);
double pawb200_get_occ(pawb200_pswf_t *wf, int band, int kpt, int spin);
double *pawb200_get_occs(pawb200_pswf_t *wf);          /* malloc'd, b*NK+kappa; free with pawb200_free_ptr */
void pawb200_set_num_sites(pawb200_pswf_t *wf, int nsites);
void pawb200_free_ptr(void *ptr);

/* ---- PAW setup (projector.h:17-20, 98-99) -------------------------------------------- */
/* labels = [label, n_channels, n_proj_grid, n_wave_grid] per element (pawpyc.pyx:368-389). */
pawb200_ppot_t *pawb200_get_projector_list(int num_els, const int *labels, const int *ls,
                                           const double *wave_grids, const double *projectors,
                                           const double *aewaves, const double *pswaves,
                                           const double *rmaxs, double grid_encut);
void pawb200_free_ppot_list(pawb200_ppot_t *pps, int length);                 /* utils.h:248 */
/* Computes <p_i|psi~_nk> for every band, k-point and spin of wf on the GPU and keeps them
 * in HBM.  Takes ownership of `pps` exactly like the reference (projector.c:570, utils.c:291). */
void pawb200_setup_projections(pawb200_pswf_t *wf, pawb200_ppot_t *pps, int num_elems,
                               int num_sites, const int *fftg, const int *labels,
                               const double *coords);

/* ---- band-pair overlaps (pseudoprojector.h:28-29; projector.h:109-128) --------------- */
void pawb200_pseudoprojection(pawb200_c128 *projections, pawb200_pswf_t *wf_ref,
                              pawb200_pswf_t *wf_proj, int BAND_NUM, int flip_spin);
void pawb200_overlap_setup_real(pawb200_pswf_t *wf_R, pawb200_pswf_t *wf_S,
                                const int *labels_R, const int *labels_S,
                                const double *coords_R, const double *coords_S,
                                const int *N_R, const int *N_S, const int *N_RS_R,
                                const int *N_RS_S, int num_N_R, int num_N_S, int num_N_RS);
void pawb200_compensation_terms(pawb200_c128 *overlap, int BAND_NUM, pawb200_pswf_t *wf_S,
                                pawb200_pswf_t *wf_R, int num_M, int num_N_R, int num_N_S,
                                int num_N_RS, const int *M_R, const int *M_S, const int *N_R,
                                const int *N_S, const int *N_RS_R, const int *N_RS_S,
                                const int *proj_labels, const double *proj_coords,
                                const int *ref_labels, const double *ref_coords,
                                const int *fft_grid, int spin_flip);

/* ---- method "aug_recip" (projector.h:114-121, 130-137; projector.c:727-848, 965-1077), SURVEY 8 row f3.
 * Same arguments as the _real pair.  The (phi - phit) augmentation of the unmatched sites is put on the FFT
 * grid, transformed forward and kept as complex64 plane-wave coefficients in HBM; compensation_terms_recip
 * accumulates (+=) <CA_R|C_S> + <C_R|CA_S> + O_M + O_N.  The plane-wave dot products accumulate in FP64
 * (the reference: single-precision cblas_cdotc_sub, projector.c:1016, 1025). */
void pawb200_overlap_setup_recip(pawb200_pswf_t *wf_R, pawb200_pswf_t *wf_S,
                                 const int *labels_R, const int *labels_S,
                                 const double *coords_R, const double *coords_S,
                                 const int *N_R, const int *N_S, const int *N_RS_R,
                                 const int *N_RS_S, int num_N_R, int num_N_S, int num_N_RS);
void pawb200_compensation_terms_recip(pawb200_c128 *overlap, int BAND_NUM, pawb200_pswf_t *wf_S,
                                      pawb200_pswf_t *wf_R, int num_M, int num_N_R, int num_N_S,
                                      int num_N_RS, const int *M_R, const int *M_S,
                                      const int *N_R, const int *N_S, const int *N_RS_R,
                                      const int *N_RS_S, const int *proj_labels,
                                      const double *proj_coords, const int *ref_labels,
                                      const double *ref_coords, const int *fft_grid, int spin_flip);

/* ---- real-space states / densities (density.h:15-67) --------------------------------- */
void pawb200_realspace_state(pawb200_c128 *x, int BAND_NUM, int KPOINT_NUM, pawb200_pswf_t *wf,
                             const int *fftg, const int *labels, const double *coords);
void pawb200_ncl_realspace_state(pawb200_c128 *x, int BAND_NUM, int KPOINT_NUM,
                                 pawb200_pswf_t *wf, const int *fftg, const int *labels,
                                 const double *coords);
void pawb200_remove_phase(pawb200_c128 *x, int KPOINT_NUM, pawb200_pswf_t *wf, const int *fftg);
void pawb200_ae_state_density(double *P, int BAND_NUM, int KPOINT_NUM, pawb200_pswf_t *wf,
                              const int *fftg, const int *labels, const double *coords);
void pawb200_ae_chg_density(double *P, pawb200_pswf_t *wf, const int *fftg, const int *labels,
                            const double *coords);
void pawb200_ncl_ae_chg_density(double *P, pawb200_pswf_t *wf, const int *fftg,
                                const int *labels, const double *coords);
void pawb200_write_volumetric(const char *filename, const double *x, const int *fftg,
                              double scale);
/* Projector method 'realspace' (density.h: project_realspace_state, density.c:205-230): band BAND_NUM of wf
 * against every band of wf_R by brute-force integration of the AE states on the grid; projs[b*NK + kappa]. */
void pawb200_project_realspace_state(pawb200_c128 *projs, int BAND_NUM, pawb200_pswf_t *wf,
                                     pawb200_pswf_t *wf_R, const int *fftg, const int *labels,
                                     const double *coords, const int *labels_R, const double *coords_R);

/* ---- FFT box (linalg.h:20-24) - one band, host buffers, GPU transform ----------------- */
void pawb200_fft3d(pawb200_c128 *x, const int *G_bounds, const double *lattice,
                   const double *kpt, const int *Gs, const pawb200_c64 *Cs, int num_waves,
                   const int *fftg);
void pawb200_fwd_fft3d(pawb200_c128 *x, const int *G_bounds, const double *lattice,
                       const double *kpt, const int *Gs, pawb200_c64 *Cs, int num_waves,
                       const int *fftg);

/* ---- small utilities the shim also binds (utils.h:204-385; pawpyc.pyx:77-149) --------- */
double pawb200_legendre(int l, int m, double x);
void pawb200_Ylm(int l, int m, double theta, double phi, double *re_im);    /* utils.h:293 */
void pawb200_Ylm2(int l, int m, double costheta, double phi, double *re_im);/* utils.h:298 */
void pawb200_frac_to_cartesian(double *coord, const double *lattice);
void pawb200_cartesian_to_frac(double *coord, const double *reclattice);
/* returns 3 rows of N doubles in one malloc'd block [3*N]; free with pawb200_free_ptr. */
double *pawb200_spline_coeff(const double *x, const double *y, int N);
double pawb200_proj_interpolate(double r, double rmax, int size, const double *x,
                                const double *proj, const double *spline3N);
double pawb200_wave_interpolate(double r, int size, const double *x, const double *f,
                                const double *spline3N);
double pawb200_spline_integral(const double *x, const double *a, const double *spline3N, int size);
/* NumSBT (sbt.h:38-58) folded into one call: k grid and transform of r*f(r), both length N */
void pawb200_spherical_bessel_transform(double encut, int l, int N, const double *r,
                                        const double *f, double *k_out, double *fk_out);
void pawb200_reciprocal_offsite_wave_overlap(const double *dcoord, const double *k1,
                                             const double *f1, const double *s1_3N, int size1,
                                             const double *k2, const double *f2,
                                             const double *s2_3N, int size2, int l1, int m1,
                                             int l2, int m2, double *re_im);   /* radial.h:30-35 */

/* ---- MomentumMatrix (momentum.h:56-92; pawpyc.pyx:738-807), SURVEY 8 row f4 -------------------------
 * < b1,k1,s1 | exp(i (G + k1 - k2).r) | b2,k2,s2 > for every G of a cutoff sphere, and the plane-wave expansion
 * of an all-electron band.  Grid helpers and quick_overlap run on the host like the reference's; the per-G sums run
 * on the GPU (one warp per G for the plane-wave correlation, one CTA per G for the one-centre terms). */
void pawb200_momentum_grid_size(pawb200_pswf_t *wf, double *nb1max, double *nb2max, double *nb3max,
                                int *npmax, double encut);
int pawb200_get_momentum_grid(int *igall, pawb200_pswf_t *wf, double nb1max, double nb2max,
                              double nb3max, double encut);
void pawb200_grid_bounds(int *G_bounds, int *gdim, const int *igall, int num_waves);
void pawb200_list_to_grid_map(int *grid, const int *G_bounds, const int *gdim, const int *igall,
                              int num_waves);
pawb200_density_ft_t *pawb200_get_all_transforms(pawb200_pswf_t *wf, double encut);
void pawb200_free_density_ft_elem_list(pawb200_density_ft_t *elems, int num_elems);
void pawb200_get_momentum_matrix(pawb200_c128 *matrix, int numg, const int *igall, pawb200_pswf_t *wf,
                                 const int *labels, const double *coords, int band1, int kpt1,
                                 int spin1, int band2, int kpt2, int spin2,
                                 pawb200_density_ft_t *transforms_list, double encut);
void pawb200_fullwf_reciprocal(pawb200_c128 *Cs, const int *igall, pawb200_pswf_t *wf, int numg,
                               int band_num, int kpt_num, const int *labels, const double *coords);
/* returns the complex value through re_im[2] (the reference returns double complex by value) */
void pawb200_quick_overlap(const int *dG, const pawb200_c128 *C1s, const pawb200_c128 *C2s, int numg,
                           const int *Gs, const int *gmap, const int *G_bounds, const int *gdim,
                           double *re_im);

/* ---- extensions (no reference counterpart; used by bench.py / batched callers) -------- */
/* Whole PAW-corrected overlap block in one call: out[kappa][b_S][b_R] (complex128,
 * nkappa*nband_S*nband_R), i.e. row b_S of block kappa is what single_band_projection(b_S)
 * returns for that kappa.  kappa_lo/kappa_hi select a shard (multi-GPU: one rank per shard).
 * pseudo_only: 0 = pseudo + aug_real augmentation, 1 = pseudo overlap only, 2 = pseudo + aug_recip. */
void pawb200_projection_matrix(pawb200_c128 *out, pawb200_pswf_t *wf_S, pawb200_pswf_t *wf_R,
                               int num_M, int num_N_R, int num_N_S, int num_N_RS,
                               const int *M_R, const int *M_S, const int *N_R, const int *N_S,
                               const int *N_RS_R, const int *N_RS_S, int flip_spin,
                               int kappa_lo, int kappa_hi, int pseudo_only);
/* Same blocks written to DEVICE memory owned by the caller, out_dev[kappa - kappa_lo][b_S][b_R] (complex128).
 * Kernels are queued on the library's main stream (the CUDA legacy default stream) and the call returns without
 * synchronising: a collective the caller launches on that stream afterwards - the NCCL all-gather of the per-k
 * matrices - needs no host round trip.  Blocks not resident on this rank are zero-filled. */
void pawb200_projection_matrix_dev(void *out_dev, pawb200_pswf_t *wf_S, pawb200_pswf_t *wf_R,
                                   int num_M, int num_N_R, int num_N_S, int num_N_RS,
                                   const int *M_R, const int *M_S, const int *N_R, const int *N_S,
                                   const int *N_RS_R, const int *N_RS_S, int flip_spin,
                                   int kappa_lo, int kappa_hi, int pseudo_only);
/* Band shard of pawb200_ae_chg_density: occupied bands in [band_lo, band_hi) only, same weights; += into P.
 * Ranks that hold the same wavefunction split the bands and all-reduce their grids (distributed.py). */
void pawb200_ae_chg_density_bands(double *P, pawb200_pswf_t *wf, const int *fftg, const int *labels,
                                  const double *coords, int band_lo, int band_hi);
/* Page-locked host memory for result buffers (e.g. the `out` of pawb200_projection_matrix): the device->host
 * copy into it runs at full PCIe rate.  Any host pointer is accepted everywhere; this is only faster. */
void *pawb200_alloc_pinned(size_t bytes);
void pawb200_free_pinned(void *p);
/* Multi-GPU sharding (one process per GPU): subsequent pawb200_read_wavefunctions* calls keep
 * only the (k,spin) blocks with kappa % world == rank in HBM; blocks of other ranks come back
 * as zeros from pawb200_projection_matrix and are summed/gathered by the caller (NCCL). */
void pawb200_set_read_shard(int rank, int world);
/* Band-block sharding of ONE (k,spin) block over ranks (SURVEY 8e level 2, for jobs with fewer (k,spin) blocks
 * than GPUs, e.g. Gamma-only): wavefunctions read after this call keep only the coefficient records of bands
 * [rank*per, (rank+1)*per), per = ceil(nband/world), and setup_projections / overlap_setup_real transform and
 * project only those bands.  Row buffers are allocated with per*world rows so the callers can all-gather the
 * blocks in place: the basis needs all its rows (coefficients, projections, wave projections) before
 * pawb200_projection_matrix[_dev], which then fills only the rows of the wf bands this rank owns (others zero).
 * Independent of pawb200_set_read_shard (which distributes whole (k,spin) blocks). */
void pawb200_set_band_shard(int rank, int world);
/* Device pointer of a row buffer of block kappa for in-place collectives.  which: 0 = plane-wave coefficients
 * (complex64 [rows][ld], box-ordered columns), 1 = projections, 2 = wave projections (complex128 [rows][ld]).
 * rows = allocated rows (per*world, x2 for spinors), [own_lo, own_hi) = the rows this rank computed.  For
 * which == 0 the library's main stream is made to wait for the coefficient copies first. */
void *pawb200_get_device_buffer(pawb200_pswf_t *wf, int which, int kappa, long *ld, int *rows, int *own_lo,
                                int *own_hi);
/* With on != 0, pawb200_read_wavefunctions_from_str returns while the host->device copies are still in
 * flight: the CALLER MUST KEEP THE BUFFER ALIVE AND UNMODIFIED until the wavefunction has been used in a call
 * that returns results (or is freed). Later launches wait per band chunk, so the transfer overlaps the
 * transforms. Only the first (k,spin) block is queued at read time; the copies of the other blocks are queued when a
 * second wavefunction is read or a consumer asks for coefficients, in block order across the wavefunctions, and
 * pawb200_overlap_setup_real queues the pseudo-overlap GEMMs of the pair on their own stream ahead of the projection
 * work. Default off (the copy is complete on return, like the reference reader). */
void pawb200_set_async_ingest(int on);
/* OpenMP threads for the host-side setup (sphere geometry, NumSBT); launchers such as torchrun export
 * OMP_NUM_THREADS=1, which would serialise it. */
void pawb200_set_host_threads(int n);
/* Drop the (k,spin) blocks outside [kappa_lo, kappa_hi) from this process. */
void pawb200_set_kappa_range(pawb200_pswf_t *wf, int kappa_lo, int kappa_hi);
/* Copy <p_i|psi~> for (band, kappa) to host: nproj_total complex128, site-major channel order.
 * which: 0 projections, 1 up_projections, 2 down_projections, 3 wave_projections. */
int pawb200_get_projections(pawb200_pswf_t *wf, int band, int kappa, int which, pawb200_c128 *out);
int pawb200_num_projections(pawb200_pswf_t *wf, int which);
/* Plane-wave coefficients of (band, kappa) in WAVECAR file order; returns the count (test accessor). */
int pawb200_get_coefficients(pawb200_pswf_t *wf, int band, int kappa, pawb200_c64 *out);
/* k-point (fractional) and weight of block kappa; returns its plane-wave count. */
int pawb200_get_kpoint(pawb200_pswf_t *wf, int kappa, double *k3, double *weight);
/* (site, n, l, m) per channel, int32[4*nproj_total] - the bit-exact index contract. */
int pawb200_get_channel_index(pawb200_pswf_t *wf, int *out);
/* sphere index list of one site as built by setup_projections (grid linear indices). */
int pawb200_get_site_indices(pawb200_pswf_t *wf, int site, int *out, int capacity);
/* per-stage device timings (ms) of the most recent call, and launch count since reset */
typedef struct {
  double h2d_ms, scatter_ms, fft_ms, project_ms, table_ms, gemm_pseudo_ms, gemm_aug_ms,
         augment_ms, d2h_ms;
  long long launches;        /* kernels of this library launched since the last reset */
  long long boxes_scattered; /* FFT boxes filled by scatter_pw_kernel */
  long long boxes_fft;       /* FFT boxes transformed */
  long long slots_projected; /* (band, table-set) sphere projections */
  long long sphere_samples;  /* psi~ samples gathered by the projection kernels (slots x sphere points) */
} pawb200_timers;
void pawb200_get_timers(pawb200_timers *t);
void pawb200_reset_timers(void);

#ifdef __cplusplus
}
#endif
#endif
