/* C ABI of the B200-native ptychography operator library (libptychofft_b200.so).
 *
 * Drop-in boundary for the compiled extension of nikitinvv/libtike-cufft: the
 * entry points below are what the reference's SWIG / pybind11 wrappers bind
 * (reference paths relative to /root/reference):
 *
 *   class ptychofft ................ src/include/ptychofft.cuh:6-44
 *   ptychofft::ptychofft ........... src/cuda/ptychofft.cu:5-41   -> ptx_create
 *   ptychofft::free / ~ptychofft ... src/cuda/ptychofft.cu:44-57  -> ptx_free / ptx_destroy
 *   ptychofft::fwd ................. src/cuda/ptychofft.cu:60-73  -> ptx_fwd
 *   ptychofft::adj (flg 0 | 1) ..... src/cuda/ptychofft.cu:76-88  -> ptx_adj
 *   read-only attrs ................ src/cuda/pybind11/ptychofft.cxx:17-22 -> ptx_dim
 *
 * and the ptx_cg_* / ptx_vec_* entry points replace the CuPy elementwise and
 * reduction code of CGPtychoSolver.run (src/libtike/cufft/ptycho.py:283-488),
 * each citing the statements it fuses.
 *
 * Conventions
 *   - plain pointers and sizes only; every array is CALLER-OWNED device memory
 *     (complex64 = interleaved float pairs, row-major, reference layouts of
 *     SURVEY.md appendix A); `stream` is a cudaStream_t passed as void*
 *     (NULL = legacy default stream, which is what the reference uses).
 *   - every function returns PTX_OK (0) or a negative PTX_E* code and records a
 *     message retrievable with ptx_last_error() (the reference checks nothing,
 *     SURVEY.md Q12).  Work is enqueued asynchronously; no host sync inside.
 *   - outputs of ptx_adj ACCUMULATE into caller-zeroed arrays, exactly like the
 *     reference's atomics (ptycho.py:102, 118); ptx_fwd overwrites g fully.
 *   - a plan owns per-CTA scratch (staging frames, accumulators, running sums) that its kernels index
 *     by blockIdx alone: a plan may be used from ONE stream at a time.  Calls on the same stream are
 *     ordered by the stream; to overlap work, overlap copies, or create one plan per stream.
 *   - supported detector sizes in this build: ndet in {64, 128, 256, 512}; nprb <= ndet.
 *     Anything else fails loudly with PTX_EUNSUPPORTED (there is no CPU or
 *     library fallback).
 */
#ifndef PTYCHOFFT_B200_H_
#define PTYCHOFFT_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTX_OK 0
#define PTX_EINVAL (-1)       /* bad argument */
#define PTX_EUNSUPPORTED (-2) /* size not built for sm_100a kernels */
#define PTX_ECUDA (-3)        /* CUDA runtime error, see ptx_last_error() */
#define PTX_EFREED (-4)       /* plan used after ptx_free() */

typedef struct ptx_plan ptx_plan;

/* noise models of CGPtychoSolver.run(model=...) (ptycho.py:308-314) */
#define PTX_MODEL_GAUSSIAN 0
#define PTX_MODEL_POISSON 1

/* ptx_dim selectors, same names as the reference's read-only attributes */
#define PTX_DIM_PTHETA 0
#define PTX_DIM_NZ 1
#define PTX_DIM_N 2
#define PTX_DIM_NSCAN 3
#define PTX_DIM_NDET 4
#define PTX_DIM_NPRB 5

const char* ptx_last_error(void);
/* 1 if a CUDA device of compute capability 10.x is present, else 0 (never throws). */
int ptx_device_ok(void);
/* number of kernels this library has launched since it was loaded (bench.py's gpu_launches). */
unsigned long long ptx_launch_count(void);

/* ptychofft::ptychofft(ptheta, nz, n, nscan, ndet, nprb), ptychofft.cu:5-41.  Allocates the
 * twiddle tables and the per-CTA scratch; no cuFFT plan, no [T,S,N,N] fft_out buffer. */
int ptx_create(ptx_plan** out, size_t ptheta, size_t nz, size_t n, size_t nscan, size_t ndet,
               size_t nprb);
int ptx_free(ptx_plan* p);    /* ptychofft::free(), idempotent (ptychofft.cu:49-57) */
int ptx_destroy(ptx_plan* p); /* ~ptychofft() */
size_t ptx_dim(const ptx_plan* p, int which);

/* g[T,S,N,N] = FFT2(pad(1/N * prb * bilinear_patch(f, scan)))     (ptychofft.cu:60-73)
 * f[T,nz,n] c64, scan[T,S,2] f32 (row, col), prb[T,P,P] c64.
 * prb_angle_stride: complex elements between consecutive angles of prb (P*P for the
 * reference layout; nmodes*P*P to address probe[:, k] of a [T,M,P,P] array in place, Q10). */
int ptx_fwd(ptx_plan* p, void* g, const void* f, const void* scan, const void* prb,
            size_t prb_angle_stride, void* stream);

/* Parity hook: near[T,S,N,N] = pad(1/N * prb * bilinear_patch(f, scan)), i.e. the output of the
 * reference's muloperator flg 2 (kernels.cu:95-107) before the FFT.  The integer work of the path
 * (patch origin, window offset (N-P)/2, skip rule) is compared bit-exactly through it. */
int ptx_debug_nearplane(ptx_plan* p, void* near, const void* f, const void* scan, const void* prb,
                        size_t prb_angle_stride, void* stream);

/* flg 0: f[T,nz,n]   += sum_s scatter(1/N * conj(prb) * crop(IFFT2(g)))   (kernels.cu:69-81)
 * flg 1: prb[T,P,P]  += sum_s 1/N * crop(IFFT2(g)) * conj(patch(f))       (kernels.cu:82-94)
 * Argument order is the reference's adj(f, g, scan, prb, flg); g is not modified. */
int ptx_adj(ptx_plan* p, void* f, const void* g, const void* scan, void* prb,
            size_t prb_angle_stride, int flg, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused CG passes.  probe is the solver's [T, M, P, P] array; scalars live in device memory.
 * ------------------------------------------------------------------------------------- */

/* ptycho.py:330-343 (and 424-428 for the probe sub-problem):
 *   I = sum_k |fwd(psi, probe[:,k])|^2 ;  red[0] += sum sqrt(I*data) ; red[1] += sum I ;
 *   red[2] += minf(I * iscale) with iscale = *iscale_dev (NULL = 1)   (ptycho.py:308-314)
 * inten_out (nullable): I is stored [T,S,N,N] f32 for the multi-mode passes.
 * red: 3 doubles, caller-zeroed. */
int ptx_cg_intensity(ptx_plan* p, const void* psi, const void* scan, const void* probe, int nmodes,
                     const float* data, float* inten_out, const float* iscale_dev, int model,
                     double* red, void* stream);

/* ptycho.py:347-363 (what = 0) and 421-441 (what = 1): one mode's gradient contribution
 *   F = fwd(psi, probe[:,mode]) ; I = inten_in*iscale (or |F|^2 when inten_in == NULL)
 *   r = F*fscale * (1 - sqrt(data)/(sqrt(I)+1e-32))      gaussian
 *   r = F*fscale * (1 - data/(I+1e-32))                  poisson
 *   what 0: grad_out[T,nz,n] += gscale * adj(r, scan, probe[:,mode])
 *   what 1: grad_out[t*grad_angle_stride + (P,P)] += gscale * adj_probe(r, scan, psi)
 * sc: 3 device floats {fscale, iscale, gscale}.  grad_out accumulates (caller zeroes it).
 * far_out (nullable): [T,S,N,N] c64, receives F itself (before the residual), for the line search
 * that follows -- re-reading 8 N^2 bytes is cheaper than gathering and transforming again. */
int ptx_cg_grad(ptx_plan* p, int what, const void* psi, const void* scan, const void* probe,
                int nmodes, int mode, const float* data, const float* inten_in, const float* sc,
                int model, void* grad_out, size_t grad_angle_stride, void* far_out, void* stream);

/* ptycho.py:383-393 (object, npairs = nmodes) and 451-461 (probe, npairs = 1):
 *   for each pair j: t1 = fwd(obj_a, prb_a[:,ja]) ; t2 = fwd(obj_b, prb_b[:,jb])
 *   p1 += |t1|^2 ; p2 += |t2|^2 ; p3 += 2 Re(t1 conj t2)      (p1 = p1_in when given)
 *   cost[0] += minf(p1) ; cost[1+c] += minf(p1 + g^2 p2 + g p3), g = 2^-(c0+c), c < 4
 * Pair j uses modes (mode_a0 + j, mode_b0 + j).  The kernel always evaluates four candidates per
 * pass (ncand <= 4 tells how many the caller will look at).  cost: 16 doubles, caller-zeroed.
 * far_a (nullable): [npairs][T,S,N,N] c64 -- t1 of every pair as left by ptx_cg_grad(far_out); when
 * given, obj_a / prb_a are not transformed again (skipped positions read as 0).
 * want_ab != 0: additionally cost[5+c] += sum sqrt(I_c data) and cost[10+c] += sum I_c for the five
 * intensities I_c evaluated (c = 0: p1), i.e. the a and b of ptycho.py:342-343 that the NEXT
 * iteration would compute from the accepted candidate; cost then is 16 doubles.
 * p23_out (nullable): [T,S,N,N] pairs of floats, receives (p2, p3) of every pixel, from which
 * ptx_cg_intensity_step forms the intensity after the accepted step without another transform. */
int ptx_cg_linesearch(ptx_plan* p, const void* obj_a, const void* prb_a, int nmodes_a, int mode_a0,
                      const void* obj_b, const void* prb_b, int nmodes_b, int mode_b0, int npairs,
                      const void* scan, const float* data, const float* p1_in, const void* far_a,
                      int model, int c0, int ncand, int want_ab, void* p23_out, double* cost,
                      void* stream);

/* inten[i] += step^2 * p23[i].p2 + step * p23[i].p3, i < n: sum_k |fwd(...)|^2 after moving `step`
 * along the direction the last line search explored (ptycho.py:424-428 recomputes it with M more
 * forward operators for every probe mode; |t1 + g t2|^2 = p1 + g^2 p2 + g p3 is the same number). */
int ptx_cg_intensity_step(float* inten, const void* p23, size_t n, float step, void* stream);

/* ---------------------------------------------------------------------------------------
 * Position correction (ptycho.py:163-248, called from the CG loop at ptycho.py:398-403).
 * ------------------------------------------------------------------------------------- */

/* register_translation_batch(src_image, target_image, upsample_factor, space), ptycho.py:192-248:
 * phase correlation of nimg pairs of N x N complex64 images (N = the plan's detector size) with the
 * reference's upsampled matrix-DFT refinement (_upsampled_dft_batch, ptycho.py:163-190, complex128).
 * fourier_space != 0: the inputs already are Fourier transforms (space='fourier'); else they are
 * transformed first (space='real').  shifts: [nimg][2] doubles (row, col), fully overwritten.
 * upsample_factor: integer in [1, 100].  The reference's trailing `shape[dim] == 1` loop (which
 * zeroes the shifts of a batch of ONE image) is left to the caller (the Python mirror does it). */
int ptx_register_translation(ptx_plan* p, const void* src, const void* target, size_t nimg,
                             int fourier_space, int upsample_factor, double* shifts, void* stream);

/* The fused position-correction step of CGPtychoSolver.run (ptycho.py:398-403):
 *   tmp1 = fwd(psi_a, scan, ones)[0] ; tmp2 = fwd(psi_b, scan, ones)[0]
 *   shifts = register_translation_batch(tmp1, tmp2, upsample_factor, space='fourier')
 * for the nscan positions of angle 0, without materialising either far field.  psi_a, psi_b:
 * [nz,n] complex64 (angle 0); scan: [nscan,2]; shifts: [nscan][2] doubles.  Skipped positions
 * (negative integer part) get the reference's all-zero-argmax value -dftshift/upsample_factor. */
int ptx_cg_position_shifts(ptx_plan* p, const void* psi_a, const void* psi_b, const void* scan,
                           int upsample_factor, double* shifts, void* stream);

/* ---------------------------------------------------------------------------------------
 * Data preparation in front of the solver (the reference's catalyst driver,
 * tests/catalyst/test_rec_script.py:44-46, 98-100, 209; ":43 TODO ... moved to compute kernels").
 * ------------------------------------------------------------------------------------- */

/* out[s][y][x] = raw[ids[s]][(y + n/2) % n][(x + n/2) % n] / denominator, s < nsel: frame selection
 * (ids == NULL: frame s), fftshift (when fftshift != 0) and normalisation in one HBM pass.
 * raw: [*, n, n] float32 device array, ids: nsel int64 on the device, out: [nsel, n, n]. */
int ptx_prepare_data(const float* raw, const long long* ids, size_t nsel, size_t n, float denominator,
                     int fftshift, float* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Small fused vector kernels on object / probe sized complex arrays (n complex elements),
 * replacing the CuPy temporaries of ptycho.py:344, 356, 366-372, 405, 435, 444-450, 463.
 * ------------------------------------------------------------------------------------- */

/* red[0] += sum |g|^2 ; red[1..2] += sum conj(d) * (g - g0)   (ptycho.py:369-371, 447-449) */
int ptx_vec_dai_yuan_reduce(const void* g, const void* g0, const void* d, size_t n, double* red,
                            void* stream);
/* first & 1: d = -g ; else d = -g + (red[0] / (red[1] + i red[2])) * d ; always g0 = g
 * (ptycho.py:366-372, 444-450; the complex beta of Q3 is reproduced).  first & 2: g is zeroed once it has
 * been consumed, ready for the next gradient pass to accumulate into (no separate fill launch). */
int ptx_vec_dai_yuan_update(void* g, void* g0, void* d, size_t n, const double* red,
                            int first, void* stream);
/* y += (*alpha) * x   (ptycho.py:405, 463) */
int ptx_vec_axpy(void* y, const void* x, size_t n, const float* alpha_dev, void* stream);
/* y += alpha * x, alpha passed by value (no device scalar to stage) */
int ptx_vec_axpy_s(void* y, const void* x, size_t n, float alpha, void* stream);
/* out = y + alpha * x   (psi + gammapsi * dpsi, ptycho.py:400 / 405, out of place) */
int ptx_vec_axpy_out(void* out, const void* y, const void* x, size_t n, float alpha, void* stream);
/* nbytes of x := 0 on the stream (cudaMemsetAsync; the accumulate-into outputs and the packed scalars) */
int ptx_vec_zero(void* x, size_t nbytes, void* stream);
/* scan[0, :] += shifts (ptycho.py:403): nscan (row, col) float32 pairs += float64, cast like CuPy's in-place add */
int ptx_cg_apply_shifts(float* scan, const double* shifts, size_t nscan, void* stream);
/* dst[0..2] = src[i0], src[i1], src[i2]: a, b and cost of the accepted line-search candidate become the
 * sums the next iteration opens with (ptycho.py:342-343), without a host round trip */
int ptx_cg_pick3(double* dst, const double* src, int i0, int i1, int i2, void* stream);
/* Device-side decision of one fused line-search pass, so that the host does not have to wait for the costs
 * before it queues what follows (ptycho.py:272-281, the `while f(...) > fp1` of line_search_sqr): the first of
 * the kdec (0..4; 0 accepts nothing) candidates 2^-c0, 2^-(c0+1), ... whose cost cost_row[1+j] is not above cost_row[0] is
 * accepted.  *half_step_dev = half that step (the update of ptycho.py:393, 461), or 0 when none was accepted
 * -- the _dev updates below are then no-ops; cost_row[15] = the accepted index or -1, for the host to read
 * later; carry (may be NULL) = {a, b, cost} of the intensity at the half step (cost_row[5+j+2], [10+j+2],
 * [j+2]; needs ptx_cg_linesearch(ab = 1) and kdec <= 3), or {1, 1, 0} when none was accepted. */
int ptx_cg_ls_decide(double* cost_row, int c0, int kdec, float* half_step_dev, double* carry, void* stream);
/* ptx_vec_axpy_out / ptx_cg_intensity_step with the step read from device memory (see ptx_cg_ls_decide) */
int ptx_vec_axpy_out_dev(void* out, const void* y, const void* x, size_t n, const float* alpha_dev, void* stream);
int ptx_cg_intensity_step_dev(float* inten, const void* p23, size_t n, const float* step_dev, void* stream);
/* CG scalars on the device, in the reference's float32 arithmetic (ptycho.py:342-351):
 *   red = {a, b} (doubles)  ->  *s_out = a/b ; sc[0] = fscale = b/a (gaussian) or 1 ; sc[1] = (a/b)^2 */
int ptx_cg_prep_scale(const double* red, int model, float* s_out, float* sc, void* stream);
/* sc[2] = k / (*absmax)^2   (ptycho.py:356 with k = 1; 435, 441 with k = nmodes/nscan or 1/nscan) */
int ptx_cg_prep_gscale(const float* absmax, double k, float* sc, void* stream);
/* x *= (*s)           (ptycho.py:344) */
int ptx_vec_scale(void* x, size_t n, const float* s_dev, void* stream);
/* *out = max(*out, max |x|)   (ptycho.py:356, 435); out caller-zeroed */
int ptx_vec_absmax(const void* x, size_t n, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PTYCHOFFT_B200_H_ */
