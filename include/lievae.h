/* lievae.h -- C ABI of liblievae_sm100a.so (B200 / sm_100a kernels for the lie-vae SO(3) hot path).
 *
 * The reference (pimdh/lie-vae) has no FFI: its boundary is the Python surface of
 * lie_vae/lie_tools.py, lie_vae/reparameterize.py and lie_vae/decoders.py.  Each entry point below
 * replaces the ATen op chain behind one reference function (cited as file:line relative to the
 * reference tree) and is what a maintainer binds from torch.autograd.Function wrappers (ctypes stub
 * in INTEGRATION.md; lie_vae_b200/_cabi.py is that stub in this repo).
 *
 * Conventions
 *   - All pointers are DEVICE pointers to dense row-major arrays, owned by the caller; outputs are
 *     fully overwritten.  16-byte alignment gives the 128-bit fast path; unaligned spans fall back to
 *     coalesced scalar accesses (never an error).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls only enqueue
 *     work: no allocation, no synchronisation, no global mutable state; re-entrant per stream.
 *   - Return value: 0 = success; < 0 = argument error (LV_ERR_*); > 0 = cudaError_t from the launch.
 *     lv_last_error() returns the calling thread's message for the last non-zero return.
 *   - `_f32` = float, `_f64` = double (same algorithm, all arithmetic in that type).
 *   - Row counts n are int64; n = 0 is a no-op.
 */
#ifndef LIEVAE_H_
#define LIEVAE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LV_OK 0
#define LV_ERR_ARG (-1)          /* null pointer, negative size, workspace too small */
#define LV_ERR_UNSUPPORTED (-2)  /* degree > 8, k > 64, too many channels */

int lv_version(void);               /* major*10000 + minor*100 + patch */
const char* lv_last_error(void);    /* thread-local, never NULL */
int lv_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- so(3) <-> R^3: map_to_lie_algebra lie_tools.py:17-43, map_to_lie_vector lie_tools.py:46-53 ---- */
int lv_hat_fwd_f32(const float* v /*n,3*/, float* X /*n,9*/, int64_t n, void* stream);
int lv_hat_fwd_f64(const double* v, double* X, int64_t n, void* stream);
int lv_hat_bwd_f32(const float* gX /*n,9*/, float* gv /*n,3*/, int64_t n, void* stream);
int lv_hat_bwd_f64(const double* gX, double* gv, int64_t n, void* stream);
int lv_vee_fwd_f32(const float* X /*n,9*/, float* v /*n,3*/, int64_t n, void* stream);
int lv_vee_fwd_f64(const double* X, double* v, int64_t n, void* stream);
int lv_vee_bwd_f32(const float* gv /*n,3*/, float* gX /*n,9*/, int64_t n, void* stream);
int lv_vee_bwd_f64(const double* gv, double* gX, int64_t n, void* stream);

/* ---- exponential map: rodrigues lie_tools.py:56-64.  v = 0 gives R = I (the reference gives NaN) ---- */
int lv_rodrigues_fwd_f32(const float* v /*n,3*/, float* R /*n,9*/, int64_t n, void* stream);
int lv_rodrigues_fwd_f64(const double* v, double* R, int64_t n, void* stream);
int lv_rodrigues_bwd_f32(const float* v, const float* gR /*n,9*/, float* gv /*n,3*/, int64_t n, void* stream);
int lv_rodrigues_bwd_f64(const double* v, const double* gR, double* gv, int64_t n, void* stream);

/* ---- logarithm map: log_map lie_tools.py:100-109 (batched; output is the 3x3 algebra element) ---- */
int lv_log_map_fwd_f32(const float* R /*n,9*/, float* X /*n,9*/, int64_t n, void* stream);
int lv_log_map_fwd_f64(const double* R, double* X, int64_t n, void* stream);
int lv_log_map_bwd_f32(const float* R, const float* gX, float* gR, int64_t n, void* stream);
int lv_log_map_bwd_f64(const double* R, const double* gX, double* gR, int64_t n, void* stream);

/* ---- quaternions (x,y,z,w scalar-last): quaternions_to_group_matrix lie_tools.py:183-192,
 *      group_matrix_to_quaternions lie_tools.py:112-157 (Shepperd, argmax branch, 1e-6 eps) ---- */
int lv_quat_to_mat_fwd_f32(const float* q /*n,4*/, float* R /*n,9*/, int64_t n, void* stream);
int lv_quat_to_mat_fwd_f64(const double* q, double* R, int64_t n, void* stream);
int lv_quat_to_mat_bwd_f32(const float* q, const float* gR, float* gq, int64_t n, void* stream);
int lv_quat_to_mat_bwd_f64(const double* q, const double* gR, double* gq, int64_t n, void* stream);
int lv_mat_to_quat_fwd_f32(const float* R /*n,9*/, float* q /*n,4*/, int64_t n, void* stream);
int lv_mat_to_quat_fwd_f64(const double* R, double* q, int64_t n, void* stream);
int lv_mat_to_quat_bwd_f32(const float* R, const float* gq, float* gR, int64_t n, void* stream);
int lv_mat_to_quat_bwd_f64(const double* R, const double* gq, double* gR, int64_t n, void* stream);

/* ---- ZYZ Euler angles: quaternions_to_eazyz lie_tools.py:160-175, group_matrix_to_eazyz
 *      lie_tools.py:178-180 (fused: the quaternion stays in registers) ---- */
int lv_quat_to_eazyz_fwd_f32(const float* q /*n,4*/, float* e /*n,3*/, int64_t n, void* stream);
int lv_quat_to_eazyz_fwd_f64(const double* q, double* e, int64_t n, void* stream);
int lv_quat_to_eazyz_bwd_f32(const float* q, const float* ge, float* gq, int64_t n, void* stream);
int lv_quat_to_eazyz_bwd_f64(const double* q, const double* ge, double* gq, int64_t n, void* stream);
int lv_mat_to_eazyz_fwd_f32(const float* R /*n,9*/, float* e /*n,3*/, int64_t n, void* stream);
int lv_mat_to_eazyz_fwd_f64(const double* R, double* e, int64_t n, void* stream);
int lv_mat_to_eazyz_bwd_f32(const float* R, const float* ge, float* gR, int64_t n, void* stream);
int lv_mat_to_eazyz_bwd_f64(const double* R, const double* ge, double* gR, int64_t n, void* stream);

/* ---- mean maps: s2s1rodrigues lie_tools.py:67-78 (s1 = (cos, sin)), s2s2_gram_schmidt
 *      lie_tools.py:81-89 (rows e1, e2, e1 x e2), vector_to_eazyz lie_tools.py:92-97 ---- */
int lv_s2s1_rodrigues_fwd_f32(const float* s2 /*n,3*/, const float* s1 /*n,2*/, float* R /*n,9*/, int64_t n, void* stream);
int lv_s2s1_rodrigues_fwd_f64(const double* s2, const double* s1, double* R, int64_t n, void* stream);
int lv_s2s1_rodrigues_bwd_f32(const float* s2, const float* s1, const float* gR, float* gs2, float* gs1, int64_t n, void* stream);
int lv_s2s1_rodrigues_bwd_f64(const double* s2, const double* s1, const double* gR, double* gs2, double* gs1, int64_t n, void* stream);
int lv_s2s2_gram_schmidt_fwd_f32(const float* v1 /*n,3*/, const float* v2 /*n,3*/, float* R /*n,9*/, int64_t n, void* stream);
int lv_s2s2_gram_schmidt_fwd_f64(const double* v1, const double* v2, double* R, int64_t n, void* stream);
int lv_s2s2_gram_schmidt_bwd_f32(const float* v1, const float* v2, const float* gR, float* gv1, float* gv2, int64_t n, void* stream);
int lv_s2s2_gram_schmidt_bwd_f64(const double* v1, const double* v2, const double* gR, double* gv1, double* gv2, int64_t n, void* stream);
int lv_vector_to_eazyz_fwd_f32(const float* v /*n,3*/, float* e /*n,3*/, int64_t n, void* stream);
int lv_vector_to_eazyz_fwd_f64(const double* v, double* e, int64_t n, void* stream);
int lv_vector_to_eazyz_bwd_f32(const float* v, const float* ge, float* gv, int64_t n, void* stream);
int lv_vector_to_eazyz_bwd_f64(const double* v, const double* ge, double* gv, int64_t n, void* stream);

/* ---- SO(2)-subgroup equivariance distance of EquivarianceLoss.forward, losses/equivariance_loss.py:27-36:
 *      diff[i] = || Rx(theta[i]) R[i] - R2[i] ||_F^2, Rx = s2s1rodrigues(e_x, (cos, sin)); resid = Rx R - R2 is kept for the
 *      backward, which returns gR = 2 gdiff Rx^T resid and gR2 = -2 gdiff resid (theta is a random draw: no gradient) ---- */
int lv_equivariance_sqdist_fwd_f32(const float* theta /*n*/, const float* R /*n,9*/, const float* R2 /*n,9*/, float* diff /*n*/, float* resid /*n,9*/, int64_t n, void* stream);
int lv_equivariance_sqdist_fwd_f64(const double* theta, const double* R, const double* R2, double* diff, double* resid, int64_t n, void* stream);
int lv_equivariance_sqdist_bwd_f32(const float* theta, const float* resid, const float* gdiff /*n*/, float* gR, float* gR2, int64_t n, void* stream);
int lv_equivariance_sqdist_bwd_f64(const double* theta, const double* resid, const double* gdiff, double* gR, double* gR2, int64_t n, void* stream);

/* ---- out[j] = sum_i in[i*inner + j]: reduces per-sample gradients over the n axis ---- */
int lv_sum_leading_f32(const float* in, float* out, int64_t n, int64_t inner, void* stream);
int lv_sum_leading_f64(const double* in, double* out, int64_t n, int64_t inner, void* stream);

/* ---- log-sum-exp over the leading axis: utils.logsumexp utils.py:4-26 as used by VAE.log_likelihood
 *   experiments/vae.py:164-171 (importance weights over the n samples).  in (n, inner) -> out (inner);
 *   backward: gin (n, inner) = gout[j] * exp(in[i,j] - out[j]). ---- */
int lv_logsumexp_leading_fwd_f32(const float* in, float* out, int64_t n, int64_t inner, void* stream);
int lv_logsumexp_leading_fwd_f64(const double* in, double* out, int64_t n, int64_t inner, void* stream);
int lv_logsumexp_leading_bwd_f32(const float* in, const float* out, const float* gout, float* gin, int64_t n, int64_t inner,
                                 void* stream);
int lv_logsumexp_leading_bwd_f64(const double* in, const double* out, const double* gout, double* gin, int64_t n,
                                 int64_t inner, void* stream);

/* ---- fused SO(3) reparameterize + wrapped log-density.
 *   N0reparameterize.nsample reparameterize.py:137-141 (v = eps*sigma, eps explicit),
 *   SO3reparameterize.nsample reparameterize.py:269-273 (z = mu @ rodrigues(v)),
 *   SO3reparameterize.log_posterior reparameterize.py:233-263 + utils.logsumexp utils.py:4-26.
 *   mu (B,9), sigma (B,3) broadcast over n; eps (n,B,3); z (n,B,9); log_q (n,B) or NULL to skip.
 *   Backward: gz (n,B,9) or NULL (=0), glq (n,B) or NULL (=0); writes PER-SAMPLE gradients
 *   gmu (n,B,9), gsigma (n,B,3) -- for n > 1 reduce with lv_sum_leading_f32.
 *   Any 4-byte (f64: 8-byte) aligned pointers are accepted; when every tensor is 16-byte aligned full tiles move with
 *   TMA bulk copies and large launches (n == 1, or B a multiple of 256 [f64: 128]) run the persistent pipelined
 *   kernels -- same results to rounding, about 1.3x faster. ---- */
int lv_so3_reparam_fwd_f32(const float* mu, const float* sigma, const float* eps, float* z, float* log_q,
                           int64_t n, int64_t B, int k, void* stream);
int lv_so3_reparam_bwd_f32(const float* mu, const float* sigma, const float* eps, const float* gz, const float* glq,
                           float* gmu, float* gsigma, int64_t n, int64_t B, int k, void* stream);

/* ---- the same, fused with group_matrix_to_eazyz (lie_tools.py:178-180, the pose handed to the action decoder in
 *   VAE.decode, experiments/vae.py:182): angles (n,B,3) = ZYZ Euler angles of z.  z may be NULL (not stored).
 *   Backward: gangles (n,B,3) required; gz, glq may be NULL (=0). ---- */
int lv_so3_reparam_eazyz_fwd_f32(const float* mu, const float* sigma, const float* eps, float* z, float* angles,
                                 float* log_q, int64_t n, int64_t B, int k, void* stream);
int lv_so3_reparam_eazyz_bwd_f32(const float* mu, const float* sigma, const float* eps, const float* gz,
                                 const float* gangles, const float* glq, float* gmu, float* gsigma, int64_t n, int64_t B,
                                 int k, void* stream);

/* float64 instantiations of the four entry points above (same contract; the winding terms use the library exp) */
int lv_so3_reparam_fwd_f64(const double* mu, const double* sigma, const double* eps, double* z, double* log_q,
                           int64_t n, int64_t B, int k, void* stream);
int lv_so3_reparam_bwd_f64(const double* mu, const double* sigma, const double* eps, const double* gz, const double* glq,
                           double* gmu, double* gsigma, int64_t n, int64_t B, int k, void* stream);
int lv_so3_reparam_eazyz_fwd_f64(const double* mu, const double* sigma, const double* eps, double* z, double* angles,
                                 double* log_q, int64_t n, int64_t B, int k, void* stream);
int lv_so3_reparam_eazyz_bwd_f64(const double* mu, const double* sigma, const double* eps, const double* gz,
                                 const double* gangles, const double* glq, double* gmu, double* gsigma, int64_t n, int64_t B,
                                 int k, void* stream);

/* ---- the same kernels with IN-KERNEL noise (SURVEY.md section 7 "Noise"): eps ~ N(0,1) is not an input but generated per
 *   sample by Philox4x32-10 (key = seed, counter = offset + flat sample index over (n,B); Box-Muller), so the reference's
 *   `Normal(0,1).sample((n,))` (reparameterize.py:137-141) never exists in memory; the backward regenerates the same
 *   numbers from (seed, offset).  angles == NULL: plain reparameterize (z required); angles != NULL: fused with
 *   group_matrix_to_eazyz (z optional).  Backward: gangles selects the fused variant the same way.
 *   lv_philox_normal_*: out (rows,3) = exactly the eps these kernels use for samples offset .. offset + rows - 1
 *   (tests; callers that want the noise itself).  The explicit-eps entry points above stay the parity path. ---- */
int lv_so3_reparam_philox_fwd_f32(const float* mu, const float* sigma, int64_t seed, int64_t offset, float* z, float* angles,
                                  float* log_q, int64_t n, int64_t B, int k, void* stream);
int lv_so3_reparam_philox_bwd_f32(const float* mu, const float* sigma, int64_t seed, int64_t offset, const float* gz,
                                  const float* gangles, const float* glq, float* gmu, float* gsigma, int64_t n, int64_t B,
                                  int k, void* stream);
int lv_so3_reparam_philox_fwd_f64(const double* mu, const double* sigma, int64_t seed, int64_t offset, double* z, double* angles,
                                  double* log_q, int64_t n, int64_t B, int k, void* stream);
int lv_so3_reparam_philox_bwd_f64(const double* mu, const double* sigma, int64_t seed, int64_t offset, const double* gz,
                                  const double* gangles, const double* glq, double* gmu, double* gsigma, int64_t n, int64_t B,
                                  int k, void* stream);
int lv_philox_normal_f32(float* out, int64_t rows, int64_t seed, int64_t offset, void* stream);
int lv_philox_normal_f64(double* out, int64_t rows, int64_t seed, int64_t offset, void* stream);

/* ---- encoder heads fused into the reparameterize kernels: from encoder features h (B,Din), Din <= 32,
 *   mu    = mean_map(Wm h + bm)                mode 0: rodrigues (AlgebraMean reparameterize.py:148-155, Dm = 3)
 *                                              mode 1: quaternions_to_group_matrix (QuaternionMean :158-164, Dm = 4)
 *                                              mode 2: s2s2_gram_schmidt in float64 (S2S2Mean :184-197, Dm = 6)
 *                                              mode 3: normalise, s2s1rodrigues (S2S1Mean :167-181, Dm = 5: rows [s2_map; s1_map])
 *   sigma = softplus(Ws h + bs)                (N0reparameterize reparameterize.py:117-121)
 *   then exactly lv_so3_reparam(_eazyz)_fwd.  Wm (Dm,Din), bm (Dm), Ws (3,Din), bs (3): the two Linear layers' parameters.
 *   mu (B,9) / sigma (B,3) are optional outputs (module attributes); give `angles` for the Euler-fused variant (z optional).
 *   Backward: gh (n,B,Din) per sample (reduce over n with lv_sum_leading_f32), gWb ((Dm+3), Din+1): rows of the mean head
 *   then of the sigma head, each = the gradient of its weight row followed by its bias gradient, summed over all samples through `workspace`
 *   (lv_so3_head_reparam_bwd_workspace_floats floats; deterministic two-pass reduction).  gmu (B,9) / gsigma (B,3), each optional:
 *   gradients that reach the mu / sigma OUTPUTS from outside the sampler (SO3reparameterize.kl(), regularisers on mu_lie or
 *   sigma); they are added to the sampler's own gradient before the mean map / softplus are pulled back. ---- */
int64_t lv_so3_head_reparam_bwd_workspace_floats(int64_t n, int64_t B, int Din, int mode);
int lv_so3_head_reparam_fwd_f32(const float* h, const float* Wm, const float* bm, const float* Ws, const float* bs,
                                const float* eps, float* mu, float* sigma, float* z, float* angles, float* log_q, int64_t n,
                                int64_t B, int Din, int mode, int k, void* stream);
int lv_so3_head_reparam_bwd_f32(const float* h, const float* Wm, const float* bm, const float* Ws, const float* bs,
                                const float* eps, const float* gz, const float* gangles, const float* glq, const float* gmu,
                                const float* gsigma, float* gh, float* gWb, float* workspace, int64_t workspace_floats, int64_t n,
                                int64_t B, int Din, int mode, int k, void* stream);

/* ---- block-diagonal Wigner-D action on a spectrum, degrees lmin..lmax (<= 8), C channels.
 *   block_wigner_matrix_multiply lie_tools.py:226-253, wigner_d_matrix lie_tools.py:211-223,
 *   _z_rot_mat lie_tools.py:195-208, ActionNet.forward decoders.py:47-56.
 *   M = (lmax+1)^2 - lmin^2.  angles (N,3); out (N,M,C).
 *   shared_spectrum != 0: spectrum is (M,C), the same for every sample (ActionNet.item_rep);
 *   otherwise (N,M,C).  Backward only: shared_spectrum = 3 (bit 1 set) ADDS the batch sum to gspectrum instead
 *   of overwriting it (micro-batched steps accumulate one gradient).  transpose != 0 applies D^T (lie_tools.py:249-250).
 *   Backward: gout (N,M,C) -> gangles (N,3) and gspectrum ((M,C) summed over N if shared, else
 *   (N,M,C)).  The shared case needs `workspace` of lv_wigner_bwd_workspace_floats(...) floats
 *   (deterministic two-pass batch reduction, no atomics). ---- */
int64_t lv_wigner_bwd_workspace_floats(int64_t N, int lmin, int lmax, int C);
int lv_wigner_apply_fwd_f32(const float* angles, const float* spectrum, float* out, int64_t N, int lmin, int lmax,
                            int C, int shared_spectrum, int transpose, void* stream);
int lv_wigner_apply_bwd_f32(const float* angles, const float* spectrum, const float* gout, float* gangles,
                            float* gspectrum, float* workspace, int64_t workspace_floats, int64_t N, int lmin,
                            int lmax, int C, int shared_spectrum, int transpose, void* stream);

/* ---- the action fused with the reconstruction term of VAE.log_likelihood (experiments/vae.py:164-171, 199-204; toy deconv):
 *   out (N) = sum_{m,c} (D(angles_i) item_rep - x[i % B])^2 for the (n, B)-flat sample index i; the action output itself is
 *   never written.  x (B, M, C).  Forward only (the IWAE bound is an evaluation metric). ---- */
int lv_wigner_recon_sse_f32(const float* angles, const float* item_rep, const float* x, float* out, int64_t N, int64_t B,
                            int lmin, int lmax, int C, int transpose, void* stream);

/* ---- the same action for ANY degree range (lmax <= lv_wigner_generic_max_degree()) and for float64: run-time loops over
 *   a caller-owned dense J table `jtable` = J_0 | J_1 | ... | J_lmax (row-major (2l+1)^2 blocks, block l at offset
 *   l(2l-1)(2l+1)/3; lie_tools.j_matrix lie_tools.py:10-14).  Covers what the unrolled kernels above do not
 *   (degrees > 8, FP64).  Backward writes PER-COLUMN results: gangle_parts (N,C,3) -- sum over C for g_angles -- and
 *   gspectrum (N,M,C) -- sum over N for a shared spectrum; the caller reduces (deterministic, no atomics). ---- */
int lv_wigner_generic_max_degree(void);
int lv_wigner_generic_fwd_f32(const float* angles, const float* spectrum, const float* jtable, float* out, int64_t N,
                              int lmin, int lmax, int C, int shared_spectrum, int transpose, void* stream);
int lv_wigner_generic_bwd_f32(const float* angles, const float* spectrum, const float* jtable, const float* gout,
                              float* gangle_parts, float* gspectrum, int64_t N, int lmin, int lmax, int C,
                              int shared_spectrum, int transpose, void* stream);
int lv_wigner_generic_fwd_f64(const double* angles, const double* spectrum, const double* jtable, double* out, int64_t N,
                              int lmin, int lmax, int C, int shared_spectrum, int transpose, void* stream);
int lv_wigner_generic_bwd_f64(const double* angles, const double* spectrum, const double* jtable, const double* gout,
                              double* gangle_parts, double* gspectrum, int64_t N, int lmin, int lmax, int C,
                              int shared_spectrum, int transpose, void* stream);

/* ---- the consumer of the Wigner action on the tensor cores: out (M, N) = A (M, K) * Bt (N, K)^T + bias.
 *   First layer of DeconvNet, ConvTranspose2d(M*C -> hidden, 4, 1, 0) on a 1x1 input (experiments/nets.py:65-66): A = the
 *   action output y (samples, 810), Bt = the weight (810, hidden, 4, 4) viewed (810, 16*hidden) and transposed,
 *   bias index = column / bias_div (bias_div = 16); first Linear of ActionNet's MLP (decoders.py:39-41): bias_div = 1.
 *   tcgen05.mma kind::tf32 (TF32 operands, FP32 accumulation in TMEM -- cuDNN's default arithmetic for the reference's FP32
 *   convolutions), weight tiles by TMA.  A: 8-byte aligned, even lda >= K; Bt: 16-byte aligned, ldb >= K a multiple of 4
 *   floats; out: ldo >= N; bias may be NULL.  A is rounded to TF32 (nearest) inside the kernel; Bt is read with its low 13
 *   mantissa bits ignored -- round it once with lv_round_tf32_f32 (in place allowed) for unbiased products. ---- */
int lv_round_tf32_f32(const float* in, float* out, int64_t n, void* stream);
int lv_gemm_tf32_f32(const float* A, int64_t lda, const float* Bt, int64_t ldb, const float* bias, int bias_div, float* out,
                     int64_t ldo, int64_t M, int N, int K, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LIEVAE_H_ */
