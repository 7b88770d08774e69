/*
 * strainer_b200 -- C ABI of the B200-native (sm_100a) straining hot path.
 *
 * The reference (hizibu7/Strainer-GAN) is pure Python and has no FFI of its own: its boundary
 * for this path is a set of Python functions.  Each entry point below is what
 * the thin Python layer (strainer-gan_b200/api.py, same names/arguments as the reference
 * functions) binds with ctypes; the comment on each group cites the reference lines whose
 * work it replaces.  INTEGRATION.md shows the binding a maintainer of the reference adds.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; every pointer is a DEVICE pointer unless the
 *    parameter name starts with h_ (host).  The caller owns every buffer, including
 *    workspaces (sizes from the *_bytes() queries); the library allocates nothing.
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *    synchronises the device, and is re-entrant per (device, stream).
 *  - return value: SG_OK (0) or a negative SgStatus; sg_last_error_string() gives the
 *    thread-local detail.  There is NO CPU fallback: on a device that is not sm_100 every
 *    compute entry point returns SG_EARCH.
 */
#ifndef STRAINER_B200_H_
#define STRAINER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum SgStatus {
  SG_OK = 0,
  SG_EINVAL = -1,   /* bad argument */
  SG_EARCH = -2,    /* device is not sm_100 / no CUDA device */
  SG_ECUDA = -3,    /* CUDA runtime / driver error, see sg_last_error_string() */
  SG_ENOINIT = -4   /* sg_init() not called for the current device */
} SgStatus;

/* comparison used by the mask / compaction kernels (reference masks: `<` everywhere except
 * `<=` in "# z_score + DBSCAN.py:325" and `>=` in "# 상위 10% 제거해서 fake image에 concate.py:247") */
typedef enum SgCmp { SG_LT = 0, SG_LE = 1, SG_GE = 2, SG_GT = 3, SG_NOT = 4 /* OR-ed in: !(v CMP thr), the
  NaN-correct complement used for the "noisy" half of divide_dataset ("#clean...py:311") */ } SgCmp;

/* interpolation rule for a quantile from two order statistics */
typedef enum SgLerp {
  SG_LERP_NUMPY = 0, /* numpy _lerp: a+(b-a)*t, b-(b-a)*(1-t) for t>=.5, no FMA ("#strainer gan.py:381") */
  SG_LERP_TORCH = 1  /* torch.lerp: fma form ("# 상위 10% 제거해서 fake image에 concate.py:246") */
} SgLerp;

/* arithmetic of the discriminator convolutions */
typedef enum SgConvMode {
  SG_CONV_BF16 = 0,   /* bf16 operands, fp32 accumulate (1 tcgen05 pass) */
  SG_CONV_BF16X3 = 1, /* fp32-parity mode: bf16 hi/lo split, 3 tcgen05 passes (hi*hi+lo*hi+hi*lo) */
  SG_CONV_FP16 = 2    /* fp16 operands and activations, fp32 accumulate, 1 pass: 11-bit significands keep the losses within
                         1e-3 of fp32 (the north_star's fp32 bar) at the bf16 mode's speed; needs |activation| < 65504.
                         accepted by sg_d64_* and sg_ae_score_tc */
} SgConvMode;

/* memory layout of a uint8 image batch */
typedef enum SgLayout {
  SG_LAYOUT_NCHW = 0, /* [N][C][H*W]: the layout ToTensor produces */
  SG_LAYOUT_NHWC = 1  /* [N][H*W][C]: the PIL / np.asarray(image) layout ToTensor reads */
} SgLayout;

/* ---- library ------------------------------------------------------------------------- */
int sg_version(void);
const char* sg_last_error_string(void);
/* Binds the library to `device`: checks compute capability 10.x, caches the SM count, resolves
 * cuTensorMapEncodeTiled through the runtime, raises the dynamic shared-memory limits. */
int sg_init(int device);
int sg_sm_count(void);
/* Kernels this library has launched in this process so far (every launch site counts; bench.py reports the difference
 * over its timed region as gpu_launches). */
long long sg_launch_count(void);

/* ---- synthetic data (counter based, identical to oracle.synth_images) --- */
int sg_synth_images(float* out, int64_t start, int64_t count, uint32_t seed, void* stream);

/* ---- uint8 pixels -> the DataLoader's fp32 tensor --------------------------------------------
 * replaces transforms.ToTensor() + transforms.Normalize(mean, std) of "#strainer gan.py:89-90, 115-116" for a
 * dataset kept as uint8: out[n][c][p] = ((float)x / 255 - mean[c]) / std[c], every step a correctly rounded fp32
 * operation = bit-identical to the host transform.  x: `count` images of `channels` (1..4) x `plane` (= H*W) bytes
 * in `layout`; h_mean / h_std: HOST arrays of `channels` floats; out: fp32 NCHW.  HBM bound (1 B in, 4 B out). */
int sg_u8_normalize(const uint8_t* x, int64_t count, int channels, int64_t plane, int layout, const float* h_mean,
                    const float* h_std, float* out, void* stream);

/* ---- host-packed PCIe copy of a fp32 host dataset ----------------------------------------------
 * The DataLoader batches of "#strainer gan.py:364-375" are fp32 in host memory; the fp16 conv mode rounds every input
 * pixel to fp16 (RN) in conv1.  sg_host_f32_to_f16 does that rounding on the HOST (h_src, h_dst: host pointers; `threads`
 * host threads, <= 0: all the process may run on; isa 0 = best available, 1 scalar, 2 AVX2+F16C, 3 AVX-512F) so that half
 * the bytes cross PCIe; sg_f16_expand widens the fp16 bits to fp32 on the device (exact), and the scores are
 * bit-identical to the ones of the fp32 copy.  Data movement only: nothing is scored on the host.
 * sg_host_threads: CPUs in the calling process' affinity mask.  Neither host function needs sg_init. */
int sg_host_threads(void);
int sg_host_f32_to_f16(const float* h_src, int64_t count, uint16_t* h_dst, int threads, int isa);
int sg_f16_expand(const uint16_t* x, int64_t count, float* out, void* stream);

/* ---- D64 scoring: Discriminator.forward + BCE vs label 1 ------------------------------
 * replaces "#strainer gan.py:230-256" (5 convs, eval-mode BN, LeakyReLU .2, Sigmoid) and the
 * per-sample BCELoss(reduction='none') of ":369-375" / "#clean...py:279-285".               */
size_t sg_d64_packed_bytes(int conv_mode);
/* w1 [64,3,4,4] w2 [128,64,4,4] w3 [256,128,4,4] w4 [512,256,4,4] w5 [1,512,4,4] fp32 OIHW;
 * bnK = {gamma, beta, running_mean, running_var} of the BatchNorm2d after conv K (K=2,3,4). */
int sg_d64_pack(const float* w1, const float* w2, const float* w3, const float* w4, const float* w5,
                const float* bn2_gamma, const float* bn2_beta, const float* bn2_mean, const float* bn2_var,
                const float* bn3_gamma, const float* bn3_beta, const float* bn3_mean, const float* bn3_var,
                const float* bn4_gamma, const float* bn4_beta, const float* bn4_mean, const float* bn4_var,
                float bn_eps, int conv_mode, void* packed, void* stream);
size_t sg_d64_workspace_bytes(int64_t max_batch, int conv_mode);
/* x: fp32 NCHW [batch,3,64,64].  Writes logit[batch], prob[batch] = sigmoid(logit) and
 * loss[batch] = -max(log prob, -100) (any of the three may be NULL). */
int sg_d64_score(const float* x, int64_t batch, const void* packed, void* workspace, int conv_mode,
                 float* logit, float* prob, float* loss, void* stream);
/* Train-mode-BatchNorm scoring: what `netD(real)` computes when netD was never put in eval mode
 * ("# 상위 10% 제거해서 fake image에 concate.py:244-245"): layers 2..4 normalise with
 * the batch statistics of THIS call and update the running statistics in place
 * (running = (1-momentum)*running + momentum*batch_stat, unbiased variance); NULL running pointers
 * skip the update.  gamma/beta come from `packed`. */
int sg_d64_score_train(const float* x, int64_t batch, const void* packed, void* workspace, int conv_mode,
                       float* bn2_running_mean, float* bn2_running_var, float* bn3_running_mean,
                       float* bn3_running_var, float* bn4_running_mean, float* bn4_running_var, float momentum,
                       float bn_eps, float* logit, float* prob, float* loss, void* stream);
/* One stage of sg_d64_score on the same workspace (1: conv1 ... 4: conv4, 5: head); used by the
 * benchmark to time each kernel with events on the launching stream, and by the tests. */
int sg_d64_run_layer(const float* x, int64_t batch, const void* packed, void* workspace, int conv_mode, int layer,
                     float* logit, float* prob, float* loss, void* stream);
/* Status words.  The conv kernels bound every mbarrier wait (2 s) and record the role that gave up in word 0; in
 * SG_CONV_FP16 mode the head records a non-finite logit (an activation beyond 65504) in word 1.  Both words are
 * STICKY: they live in the first 8 bytes of `workspace` and stay set across calls until sg_d64_check reports and
 * clears them.  The *_status variants write to a caller-owned, caller-zeroed device int32[2] instead (NULL = the
 * workspace's words), so that a caller can keep one pair per chunk and re-score exactly the chunks that overflowed
 * (api.py conv_mode="auto").  In SG_CONV_FP16 mode sg_d64_score_train_status commits the running statistics only if
 * word 1 is still zero after the head: a batch that overflowed leaves them untouched. */
int sg_d64_score_status(const float* x, int64_t batch, const void* packed, void* workspace, int conv_mode,
                        float* logit, float* prob, float* loss, int32_t* status2, void* stream);
int sg_d64_score_train_status(const float* x, int64_t batch, const void* packed, void* workspace, int conv_mode,
                              float* bn2_running_mean, float* bn2_running_var, float* bn3_running_mean,
                              float* bn3_running_var, float* bn4_running_mean, float* bn4_running_var, float momentum,
                              float bn_eps, float* logit, float* prob, float* loss, int32_t* status2, void* stream);
/* Synchronises `stream`, reports (SG_ECUDA: pipeline time-out, SG_EINVAL: fp16 overflow) and clears the workspace's
 * own status words. */
int sg_d64_check(const void* workspace, void* stream);
/* debugging / tests: copies the activation of layer `layer` (1..4) of the last sg_d64_score call
 * on `workspace` into fp32 NCHW `out` ([batch,64,32,32], [batch,128,16,16], ...). */
int sg_d64_read_activation(const void* workspace, int64_t batch, int conv_mode, int layer, float* out,
                           void* stream);

/* ---- D64 training step: forward with batch-statistics BatchNorm + the whole backward pass ------------------
 * replaces what autograd runs for `output = netD(x)` and `err.backward()` in the D and G steps of
 * "#strainer gan.py:586-633" (":589-592" D on real, ":598-603" D on fake.detach(), ":610-615" G through D).
 * precision 0: fp16 operands on tcgen05, fp32 accumulation (the arithmetic class of torch's default TF32 convolutions);
 * precision 1: split operands, x = hi + lo in fp16, every GEMM as A_hi.B_hi + A_lo.B_hi + A_hi.B_lo (three tensor passes, the
 * fp32-parity arithmetic: gradients within 1e-3 of fp32 autograd).  Packed block, workspace, forward and backward of one
 * pass must use the same precision.  Gradients carry a per-call power-of-two loss scale.
 *
 * packed: sg_d64_train_packed_bytes() bytes, 1024-byte aligned: the fp16 operand forms of the five conv weights, written by
 * sg_d64_train_pack (h_weights: HOST array of 5 DEVICE pointers, conv1..conv5 weight [Cout][Cin][4][4]); repack after
 * every optimiser step.  The backward pass reads it too: it must still hold the weights its forward ran with.
 * workspace: sg_d64_train_workspace_bytes(max_batch) bytes, 1024-byte aligned, prepared ONCE by
 * sg_d64_train_workspace_init (zero borders of the padded activation / gradient tensors); it keeps everything the
 * backward pass needs, so one workspace serves one forward -> backward pair at a time.
 * h_bn_params: HOST array of 6 DEVICE pointers {bn2 gamma, bn2 beta, bn3 gamma, bn3 beta, bn4 gamma, bn4 beta};
 * h_running_stats: HOST array of 6 DEVICE pointers {bn2 mean, bn2 var, ...} updated in place as nn.BatchNorm2d does in
 * training mode (NULL: no update); they are committed at the end of the forward only if every logit of the batch is finite
 * (an fp16 overflow leaves them untouched, so that the caller can score the batch again in another arithmetic).  x fp32 NCHW [batch,3,64,64], 2 <= batch <= max_batch.
 * forward writes prob[batch] = sigmoid(logit) and logit[batch] (either may be NULL).
 * backward takes grad_prob[batch] = dL/dprob and writes h_grads (HOST array of 11 DEVICE pointers {dconv1..dconv5 weight,
 * dgamma2, dbeta2, dgamma3, dbeta3, dgamma4, dbeta4} in PyTorch layouts; NULL skips every parameter gradient, as the G
 * step may) and grad_x [batch,3,64,64] (NULL skips it). */
size_t sg_d64_train_packed_bytes(int precision);
int sg_d64_train_pack(const float* const* h_weights, int precision, void* packed, void* stream);
size_t sg_d64_train_workspace_bytes(int64_t max_batch, int precision);
int sg_d64_train_workspace_init(void* workspace, int64_t max_batch, int precision, void* stream);
int sg_d64_train_forward(const float* x, int64_t batch, int64_t max_batch, int precision, const void* packed,
                         const float* const* h_bn_params, float* const* h_running_stats, float momentum, float bn_eps,
                         void* workspace, float* prob, float* logit, void* stream);
int sg_d64_train_backward(const float* grad_prob, int64_t batch, int64_t max_batch, int precision, const void* packed,
                          void* workspace, float* const* h_grads, float* grad_x, void* stream);
/* synchronises `stream`; SG_ECUDA: a GEMM pipeline timed out, SG_EINVAL: a non-finite logit or gradient was produced
 * (fp16 range); clears the status words */
int sg_d64_train_check(void* workspace, void* stream);
/* debugging / tests: one saved tensor as fp32 NCHW.  what: 1 act1, 2..4 raw conv output of layer 2..4, 5..6 normalised
 * activation of layer 2..3, 7 normalised activation of layer 4 */
int sg_d64_train_read(const void* workspace, int64_t batch, int64_t max_batch, int precision, int what, float* out,
                      void* stream);

/* ---- auto-encoder reconstruction-error scoring -------------------------------------------
 * replaces AutoEncoder.forward "#autoencoder.py:269-291" and the per-sample
 * F.mse_loss(out, img, 'none').view(B,-1).mean(1) of ":315-316".
 * h_params: HOST array of 12 DEVICE pointers {w,b} x {enc.0, enc.2, enc.4, dec.0, dec.2, dec.4} in the
 * PyTorch layouts (Conv2d [out,in,k,k], ConvTranspose2d [in,out,k,k]).  x fp32 NCHW [batch,3,64,64];
 * err_out[batch]; recon_out (optional, [batch,3,64,64]) receives the reconstruction.  x and recon_out must be 16-byte
 * aligned (TMA boxes and vector accesses; every sample offset of an aligned tensor is: 49 152 bytes per sample), else
 * SG_EINVAL. */
#ifdef SG_AB_VARIANTS   /* plain fp32 on the CUDA cores: experiment builds only (cross-check of the tensor-core pipeline) */
size_t sg_ae_workspace_bytes(int64_t max_batch);
int sg_ae_score(const float* x, int64_t batch, const float* const* h_params, void* workspace, float* err_out,
                float* recon_out, void* stream);
#endif

/* bf16 conv mode of the same auto-encoder (BASELINE.json config 4): the two 7x7 layers (86 % of the FLOPs) as
 * implicit GEMMs on tcgen05 with bf16 operands / fp32 accumulation, bf16 NHWC activations, the small stride-2
 * layers on the CUDA cores.  Same arguments as sg_ae_score; workspace from sg_ae_bf16_workspace_bytes, 1024-byte
 * aligned.  sg_ae_bf16_check reports a timed-out pipeline (synchronises the stream). */
size_t sg_ae_bf16_workspace_bytes(int64_t max_batch);
int sg_ae_score_bf16(const float* x, int64_t batch, const float* const* h_params, void* workspace, float* err_out,
                     float* recon_out, void* stream);
/* The same tensor-core pipeline with the conv mode as an argument: SG_CONV_BF16 as above; SG_CONV_BF16X3 = fp32-parity
 * arithmetic (activations and 7x7 weights as bf16 hi | lo, three GEMM segments, ~1e-5 relative on the errors).
 * workspace from sg_ae_tc_workspace_bytes(max_batch, conv_mode). */
size_t sg_ae_tc_workspace_bytes(int64_t max_batch, int conv_mode);
int sg_ae_score_tc(const float* x, int64_t batch, const float* const* h_params, void* workspace, int conv_mode,
                   float* err_out, float* recon_out, void* stream);
/* The two halves of sg_ae_score_tc: sg_ae_pack_tc converts the weights into the kernels' layouts once (and clears the
 * workspace's status word); sg_ae_forward_tc scores a batch with already packed weights (h_params still supplies the
 * fp32 biases) -- a dataset-scale caller packs once and runs one forward per chunk on the same workspace. */
int sg_ae_pack_tc(const float* const* h_params, void* workspace, int conv_mode, void* stream);
int sg_ae_forward_tc(const float* x, int64_t batch, const float* const* h_params, void* workspace, int conv_mode,
                     float* err_out, float* recon_out, void* stream);
int sg_ae_bf16_check(const void* workspace, void* stream);

/* ---- MLP discriminator scoring (28x28 path) --------------------------------------------------
 * replaces Discriminator.forward of "Untitled-2.py:79-94" / "# 1,2,8.py:110-128" (eval mode) + BCE vs 1.
 * h_params: HOST array of 8 DEVICE pointers {weight [out,in], bias} x 4 Linear layers.  x fp32 [batch,784]. */
size_t sg_mlp_workspace_bytes(int64_t max_batch);
int sg_mlp_score(const float* x, int64_t batch, const float* const* h_params, void* workspace, float* logit,
                 float* prob, float* loss, void* stream);

/* The same chain on the tensor cores for dataset-scale batches ("tcgen05 GEMM chain"): fp32 rows -> fp16,
 * three tcgen05 GEMMs (fp16 operands, fp32 accumulation, fused bias + LeakyReLU epilogues), dot + sigmoid + BCE head.
 * sg_mlp_tc_pack converts the three hidden-layer weights to fp16 once (packed: sg_mlp_tc_packed_bytes, 1024-aligned);
 * h_params still supplies the fp32 biases and the last layer.  status2 (device int32[2], caller-zeroed, may be NULL = the
 * first words of `workspace`): [0] pipeline time-out role code, [1] non-zero if a logit came out non-finite (an
 * activation beyond fp16's range: score that batch with sg_mlp_score). */
size_t sg_mlp_tc_packed_bytes(void);
size_t sg_mlp_tc_workspace_bytes(int64_t max_batch);
int sg_mlp_tc_pack(const float* const* h_params, void* packed, void* stream);
int sg_mlp_score_tc(const float* x, int64_t batch, const float* const* h_params, const void* packed, void* workspace,
                    float* logit, float* prob, float* loss, int32_t* status2, void* stream);

/* ---- DCGAN-28 conv discriminator scoring (BASELINE.json config 1) -------------------------------
 * The reference has no 28x28 DCGAN; SURVEY 8d C1 option (ii) defines one for config 1 and this is it:
 * Conv(1->64, k4 s2 p1) LeakyReLU(.2) | Conv(64->128, k4 s2 p1) BatchNorm2d LeakyReLU(.2) | Conv(128->1, k7) Sigmoid,
 * all convs without bias; eval-mode BN folded.  Layer 1 runs on the CUDA cores and writes the im2col rows of layer 2,
 * which is one tcgen05 GEMM [batch*49, 1024] x [128, 1024]^T; layer 3 + sigmoid + BCE is a dot per image.
 * w1 [64,1,4,4] fp32 is read as is; sg_d28_pack packs w2 [128,64,4,4], w3 [1,128,7,7] and the BN fold.
 * x fp32 [batch,1,28,28].  status2 as for sg_mlp_score_tc. */
size_t sg_d28_packed_bytes(void);
size_t sg_d28_workspace_bytes(int64_t max_batch);
int sg_d28_pack(const float* w2, const float* w3, const float* bn_gamma, const float* bn_beta, const float* bn_mean,
                const float* bn_var, float bn_eps, void* packed, void* stream);
int sg_d28_score(const float* x, int64_t batch, const float* w1, const void* packed, void* workspace, float* logit,
                 float* prob, float* loss, int32_t* status2, void* stream);

/* ---- selection: order statistics, thresholds ------------------------------------------
 * replaces np.percentile "#strainer gan.py:381", "# 종합 loss.py:288-292" and torch.quantile
 * "# 상위 10% 제거해서 fake image에 concate.py:246", "# z_score + DBSCAN.py:323".
 * Radix select of the order statistics x_(k) and x_(k+1) (NaNs sort last, -0 == +0) over 4 x 8 key bits
 * (256-bin histograms in 32 lane-private shared-memory copies: conflict free even for skewed data); the last
 * pass also tracks the smallest key above the selected 24-bit bucket, so x_(k+1) costs no extra read.
 *  - single device: sg_select_kth / sg_radix_select.  All four passes run in ONE cooperative kernel (grid barrier
 *    between passes, the CTA's slice of the input cached in shared memory); for n >= SG_SELECT_ONEPASS_MIN the
 *    input itself is read ONCE (sampled pivots + a streaming filter pass) and the passes run over the ~2 % of
 *    candidates between the pivots.
 *  - multi-GPU: the phase entry points below, so that the caller can all-reduce ws[0..257) (uint32 digit
 *    histogram + NaN count, SUM) after every sg_select_hist and ws[SG_SELECT_WS_MINABOVE] (MIN; as int32 of
 *    key ^ 0x80000000 it is order preserving) after the last one, before the matching sg_select_step.  */
#define SG_SELECT_WS_WORDS 2048      /* uint32 words of workspace */
#define SG_SELECT_WS_HIST 0          /* [256] digit histogram of the current pass */
#define SG_SELECT_WS_NANCOUNT 256    /* number of NaNs seen (pass 0); SUM-reduced with the histogram */
#define SG_SELECT_WS_MINABOVE 257    /* smallest radix key above the selected bucket (MIN-reduced) */
#define SG_SELECT_NUM_PASSES 4
int sg_select_begin(uint32_t* ws, int64_t k, void* stream);
int sg_select_hist(const float* v, int64_t n, uint32_t* ws, int pass, void* stream);
int sg_select_step(uint32_t* ws, int pass, void* stream);
/* Multi-GPU without a collective launch: sg_select_step_peer = the all-reduce of pass `pass` (SUM of ws[0..257), MIN of
 * ws[257] on the last pass) FUSED into sg_select_step, over NVLink peer memory.  Every rank owns a buffer of
 * sg_peer_buffer_bytes(nranks) that all ranks have mapped; d_peer_buffers is a DEVICE array of the nranks mapped
 * addresses (own buffer at index `rank`).  One-shot push protocol: each word travels as one 64-bit store tagged with
 * `seq` (non-zero, incremented by every rank for every call, in the same order on all ranks); a rank polls only its own
 * buffer.  All ranks must enqueue the call; each rank's GPU must run it concurrently with its peers' (one process per
 * GPU).  A peer that has not arrived after 20 s is reported by sg_select_check (status 2), not by a hang (the bound covers the
 * skew between ranks whose host-to-device copies finish seconds apart).
 * sg_peer_alloc / sg_peer_open: the one allocation the library makes itself (a CUDA IPC handle needs a cudaMalloc'ed
 * base): h_handle64 is a HOST buffer of 64 bytes to be exchanged between the processes (e.g. all_gather). */
size_t sg_peer_buffer_bytes(int nranks);
int sg_peer_alloc(int nranks, void** buffer_out, void* h_handle64_out);
int sg_peer_open(const void* h_handle64, void** peer_out);
int sg_peer_close(void* peer);
int sg_peer_free(void* buffer);
int sg_select_step_peer(uint32_t* ws, int pass, void* const* d_peer_buffers, int rank, int nranks, uint32_t seq,
                        void* stream);
/* out2[0] = x_(k), out2[1] = x_(k+1) (== x_(k) when k is the last index); NaN if any NaN. */
int sg_select_finish(const uint32_t* ws, float* out2, void* stream);
/* single-device convenience: all phases back to back (each histogram pass ends with its own bucket step,
 * run by the last CTA to finish: 4 + 2 launches) */
int sg_radix_select(const float* v, int64_t n, int64_t k, uint32_t* ws, float* out2, void* stream);
/* One-pass selection for large n (>= SG_SELECT_ONEPASS_MIN): two pivots from a 32768-element sample, ONE
 * streaming read of v that counts the elements below the lower pivot and collects the ~3 % between the pivots,
 * then the radix passes over that small buffer.  Always exact: if the pivots miss x_(k)/x_(k+1) or the buffer
 * overflows (heavy ties) the radix passes run over v itself.  workspace: sg_select_workspace_bytes(n) bytes,
 * 16-byte aligned; with a smaller workspace (>= 2 KB) it degrades to sg_radix_select. */
#define SG_SELECT_ONEPASS_MIN (1 << 21)
size_t sg_select_workspace_bytes(int64_t n);
int sg_select_kth(const float* v, int64_t n, int64_t k, void* workspace, size_t workspace_bytes, float* out2,
                  void* stream);
/* Synchronises `stream` and reports (SG_ECUDA) a grid barrier of the cooperative select kernel that gave up after its
 * 2 s bound -- possible only when the device is time-sliced (MPS, debugger) under the cooperative launch; the order
 * statistics of that call were written as NaN, never as values from incomplete histograms. */
int sg_select_check(const void* workspace, void* stream);
/* thr[0] = lerp(stats2[0], stats2[1], weight) with the named rounding rule */
int sg_lerp_threshold(const float* stats2, float weight, int lerp_kind, float* thr, void* stream);
/* Segmented in-block quantile: `segments` consecutive segments of `seg_len` (<= 2048) values each;
 * per segment a shared-memory bitonic sort, out2[2*s] = x_(k), out2[2*s+1] = x_(k1). */
int sg_segment_order_stats(const float* v, int64_t segments, int seg_len, int k, int k1, float* out2,
                           void* stream);

/* ---- compaction ----------------------------------------------------------------------
 * replaces np.where(losses < thr)[0] "#strainer gan.py:384", the boolean gathers
 * real_cpu[mask] / real_cpu[~mask] and torch.cat "# 상위 10% 제거...py:247-249,268", and the pool
 * gather "# strainer gan + concate.py:623-627".                                               */
size_t sg_compact_workspace_bytes(int64_t n);
/* idx_out[0..count) = ascending i (+ index_base) with v[i] CMP *thr; *count_out = count.
 * If mask_out != NULL also writes the 0/1 byte mask. */
int sg_compact_indices(const float* v, int64_t n, const float* thr, int cmp, int64_t index_base,
                       int64_t* idx_out, int64_t* count_out, uint8_t* mask_out, void* workspace, void* stream);
/* Stable two-way partition of rows by mask: kept rows (mask != 0) to `kept`, the others to
 * `dropped` (either may be NULL); counts_out[0] = #kept, counts_out[1] = #dropped.
 * row_bytes must be a multiple of 16. */
int sg_compact_rows(const void* rows, int64_t n, int64_t row_bytes, const uint8_t* mask, void* kept,
                    void* dropped, int64_t* counts_out, void* workspace, void* stream);
/* The selection half of the in-batch strain block in two launches, for one batch of n <= 2048 scores
 * ("# 상위 10% 제거해서 fake image에 concate.py:246-249"): thr = lerp(x_(k0), x_(k1), weight) by the named rule
 * (torch.quantile), mask[i] = scores[i] CMP thr, then the stable partition rows[mask] -> kept, rows[~mask] ->
 * dropped (rows may be NULL: threshold, mask and counts only).  workspace: n int64 (8*n bytes).
 * counts_out = {#kept, #dropped}. */
int sg_strain_rows(const float* scores, int64_t n, int k0, int k1, float weight, int lerp_kind, int cmp,
                   const void* rows, int64_t row_bytes, void* kept, void* dropped, uint8_t* mask_out, float* thr_out,
                   int64_t* counts_out, void* workspace, void* stream);

/* The same with the torch.cat of ":268" fused away: `concat` is ONE buffer of n rows; the dropped rows (the strained
 * reals that join the fake batch) are written to its rows [#kept, n) -- straight behind the #kept rows the generator
 * output is then copied to by sg_concat_rows -- instead of to a buffer of their own. */
int sg_strain_rows_concat(const float* scores, int64_t n, int k0, int k1, float weight, int lerp_kind, int cmp,
                          const void* rows, int64_t row_bytes, void* kept, void* concat, uint8_t* mask_out, float* thr_out,
                          int64_t* counts_out, void* workspace, void* stream);
/* out[0..na) = a, out[na..na+nb) = b (rows of row_bytes, a multiple of 16): torch.cat([fake, filtered_fake], 0) of
 * ":268" as one launch; rows of b that already sit at out + na*row_bytes (sg_strain_rows_concat) are not touched. */
int sg_concat_rows(const void* a, int64_t na, const void* b, int64_t nb, int64_t row_bytes, void* out, void* stream);

/* out[i] = rows[idx[i]] for i < count (count read from *count_dev if non-NULL, else `count`). */
int sg_gather_rows(const void* rows, int64_t row_bytes, const int64_t* idx, int64_t count,
                   const int64_t* count_dev, void* out, void* stream);

/* ---- moments / z-score / histogram ------------------------------------------------------
 * replaces err.mean() + k*err.std() "#autoencoder.py:320", the feature z-score
 * "#z_score.py:286-291" / "# 1,2,8.py:164-168" and np.histogram "#strainer gan.py:293".        */
#define SG_MOMENT_CHUNK 4096
/* partial[2*c] = sum, partial[2*c+1] = sum of squares (fp64, fixed pairwise order) of chunk c of
 * SG_MOMENT_CHUNK values; chunks are in index order so any sharding that aligns to the chunk
 * size reproduces the single-device partials bit for bit. */
int sg_chunk_moments(const float* v, int64_t n, double* partial, void* stream);
/* stats[0] = mean, stats[1] = unbiased std (fp64) of n values from `chunks` partials, summed in
 * chunk order; thr[0] = (float)mean + k * (float)std as torch evaluates it in fp32. */
int sg_moments_finish(const double* partial, int64_t chunks, int64_t n, float k, double* stats, float* thr,
                      void* stream);
/* out[i] = (float)((v[i] - stats[0]) / std) evaluated in fp64: StandardScaler on a single column
 * ("# z_score + DBSCAN.py:291") from the stats of sg_moments_finish (stats[1] = unbiased std; ddof 0 rescales it to the
 * population std StandardScaler uses; a zero std leaves the values unscaled). */
int sg_standardize(const float* v, int64_t n, const double* stats, int ddof, float* out, void* stream);
size_t sg_col_moments_workspace_bytes(int64_t n, int d);
/* mean[d], inv-or-std[d] of the columns of x[n,d]; ddof 1 (torch.std) or 0 (np.std);
 * denom[j] = std_j + eps_add. */
int sg_col_moments(const float* x, int64_t n, int d, int ddof, float eps_add, float* mean, float* denom,
                   void* workspace, void* stream);
/* out[i] = max_j |(x[i,j] - mean[j]) / denom[j]| (NaN propagates like torch.max) */
int sg_row_max_absz(const float* x, int64_t n, int d, const float* mean, const float* denom, float* out,
                    void* stream);
/* minmax: device buffer of 8 floats; [0] = min, [1] = max (both NaN if any NaN), [2..8) scratch */
int sg_minmax(const float* v, int64_t n, float* minmax, void* stream);
/* np.histogram's uniform-bin fast path: edges[bins+1] fp32 (as np.linspace made them),
 * counts[bins] int64 accumulated (caller zeroes). */
int sg_hist_uniform(const float* v, int64_t n, const float* edges, int bins, long long* counts, void* stream);

/* ---- DBSCAN clean ratio of [n, d] features on the tensor cores -----------------------------------
 * replaces StandardScaler -> DBSCAN(eps, min_samples) -> mean(labels != -1) of estimate_ratio_dbscan
 * "# z_score + DBSCAN.py:291-299" (SURVEY 8f item 2).  z = (x - mean) / denom (column moments from sg_col_moments,
 * ddof 0), pairwise squared distances as a tcgen05 GEMM (bf16 hi/lo split, fp32-grade) with the eps threshold in
 * the epilogue: pass 0 neighbour counts -> core points, pass 1 points within eps of a core point.
 * counts_out = {#core, #non-noise} (device int64[2]); d multiple of 64; workspace 1024-byte aligned. */
size_t sg_dbscan_nd_workspace_bytes(int64_t n, int d);
int sg_dbscan_nd(const float* x, int64_t n, int d, const float* mean, const float* denom, double eps, int min_samples,
                 int64_t* counts_out, void* workspace, void* stream);
int sg_dbscan_nd_check(const void* workspace, void* stream);

/* ---- two-component 1-D Gaussian-mixture EM ----------------------------------------------------
 * replaces GaussianMixture(n_components=2, max_iter=10, tol=1e-2, reg_covar=5e-4).fit(losses) of
 * "#clean 분포와 noisy 분포가 만나는 지점의 loss보다 작은 데.py:290-292", "# 종합 loss.py:271-273" (SURVEY 8f item 2).
 * scikit-learn's EM equations; deterministic start (Lloyd iterations from two given centres) instead of the
 * RNG-seeded k-means.  Protocol: sg_gmm1d_begin, then (kmeans_iters + max_iter) rounds of
 * sg_gmm1d_accumulate -> [multi-GPU: all-reduce SUM of the 8 doubles at workspace + 128] -> sg_gmm1d_update;
 * rounds after convergence are no-ops, nothing is read back in between.  Result: doubles at workspace + 0:
 * weights[2], means[2], variances[2], lower bound, n_iter, converged. */
size_t sg_gmm1d_workspace_bytes(void);
int sg_gmm1d_begin(const float* init_centers2, int kmeans_iters, void* workspace, void* stream);
int sg_gmm1d_accumulate(const float* v, int64_t n, void* workspace, void* stream);
int sg_gmm1d_update(int64_t n_total, double reg_covar, double tol, int max_iter, void* workspace, void* stream);

/* ---- device sort + 1-D DBSCAN clean ratio --------------------------------------------------
 * BASELINE.json north_star: "1-D DBSCAN thresholds on a device sort".  The reference's
 * estimate_ratio_dbscan ("# z_score + DBSCAN.py:272-301") consumes only the fraction of
 * non-noise points; this is sklearn.cluster.DBSCAN(eps, min_samples) applied to an (N,1) array:
 * counts_out[0] = number of non-noise points, noise_out (optional) = 0/1 per point in input order. */
size_t sg_sort_workspace_bytes(int64_t n);
/* ascending stable LSD radix sort (NaN last, -0 == +0), n <= 2^30: one histogram kernel + 4 single-kernel passes
 * (decoupled look-back).  sorted_out and / or order_out (source index of each sorted element) may be NULL: without
 * order_out the passes move keys only (36 B / element instead of 64).  Outputs must not overlap the workspace. */
int sg_sort_f32(const float* v, int64_t n, float* sorted_out, int32_t* order_out, void* workspace, void* stream);
size_t sg_dbscan1d_workspace_bytes(int64_t n);
int sg_dbscan1d(const float* v, int64_t n, double eps, int min_samples, int64_t* counts_out, uint8_t* noise_out,
                void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STRAINER_B200_H_ */
