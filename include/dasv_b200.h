/* dasv_b200 — C ABI of the B200-native DoubleMHA speaker-embedding extraction path.
 *
 * The reference (fedecosta/DoubleAttentionSpeakerVerification) has no FFI of its own: its
 * boundary is the Python nn.Module API of scripts/CNNs.py, scripts/poolings.py and
 * scripts/model.py.  The Python classes of the same names in
 * doubleattentionspeakerverification_b200/ keep that API and call ONLY the entry points below
 * (through ctypes).  Each entry point cites the reference lines it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; the caller owns all memory,
 *     including outputs and workspaces (sizes are given per function);
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises the host, allocates,
 *     or creates streams; all functions are re-entrant per device;
 *   - return 0 on success; non-zero = error, message via dasv_last_error() (thread-local);
 *   - dtype codes: DASV_F32 = 0, DASV_BF16 = 1, DASV_F16 = 2 (front-end activations / packed weights only);
 *   - "nullable" arguments may be NULL to skip that input/output.
 */
#ifndef DASV_B200_H
#define DASV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DASV_F32 0
#define DASV_BF16 1
#define DASV_F16 2
#define DASV_SPLIT_BF16 3            /* conv11 output only: [hi(C) | lo(C)] bf16 pairs of fp32 values (fp32x3 mode) */

/* conv flags */
#define DASV_CONV_RELU 1        /* apply ReLU after bias (set by the VGG blocks; clear = linear, for the input-gradient pass, no POOL) */
#define DASV_CONV_POOL 2        /* fuse max_pool2d(2, stride 2, ceil_mode) into the epilogue         */
#define DASV_CONV_REF_LAYOUT 4  /* with POOL: write [B,T',C*F'] with feature = c*F'+f (CNNs.py:88-89) */
#define DASV_CONV_PAIR 8        /* run on CTA pairs (tcgen05 cta_group::2, 256 channels x N pixels per pair); same results */
#define DASV_CONV_W_F16 16      /* wp was packed as fp16 (dasv_pack_conv_weight_16 with DASV_F16) */
#define DASV_CONV_X_F16 32      /* x (and an NHWC y) are fp16 instead of bf16; fp16 stores saturate at +-65504 */
#define DASV_CONV_X3 64         /* fp32x3 mode: x is [B,T,F,2*Cin] split bf16, wp from dasv_pack_conv_weight_x3, an NHWC y is [..,2*Cout] split bf16 */
#define DASV_CONV_LAZY_MASK 256 /* with lengths: of the rows at or beyond an utterance's length L, only the one the NEXT 3x3 layer reads is
                                   guaranteed zero (row L; with POOL: pooled row ceil(L/2)); all-masked tiles further down are skipped and
                                   the memory behind them is left unwritten.  For pipelines that carry `lengths` through every layer. */

int dasv_abi_version(void);
const char* dasv_last_error(void);

/* ---------------------------------------------------------------- DoubleMHA pooling
 * Replaces DoubleMHA.forward = MultiHeadAttention.forward -> HeadAttention.forward
 * (scripts/poolings.py:126-129 -> :100-109 -> innerKeyValueAttention :73-80 -> :45-51,:61-71)
 * with ONE pass over x:
 *   s[b,t,h] = <x[b,t,h,:], query[:,h]> / sqrt(H);  p = softmax_t(s) over t < lengths[b];
 *   ctx[b,h,:] = sum_t p x;  u[b,h] = <ctx[b,h,:], att>;  u = -inf where keep==0;
 *   headw = softmax_h(u);  out[b,:] = sum_h headw ctx[b,h,:].
 * x [B,T,D] (x_dtype), D = H*dh;  lengths [B] int32 nullable (NULL = all T);
 * query [dh,H] f32 (reference layout);  att [dh] f32 nullable (NULL = MultiHeadAttention only:
 * no head stage, `out`/`headw` must then be NULL);  keep [B,H] uint8 nullable (training-mode
 * head drop-out mask, poolings.py:39-43, drawn by the caller);
 * out [B,dh], ctx [B,H,dh], lse [B,H], headw [B,H], align [B,T,H] : f32, each nullable.
 * workspace (nullable, dasv_dmha_fwd_workspace_bytes): the utterance counter of the dynamic deal (CTAs claim the
 * next utterance when they are ready for one, so ragged lengths do not leave a CTA with two long utterances);
 * without it utterances are dealt round-robin.  The caller ZEROES it once; every launch leaves it zeroed again,
 * so one buffer can be reused by consecutive calls on the same stream (not by concurrent streams).
 */
size_t dasv_dmha_fwd_workspace_bytes(int B, int T, int D, int H);
int dasv_dmha_fwd(const void* x, int x_dtype, const int32_t* lengths,
                  const float* query, const float* att, const uint8_t* keep,
                  float* out, float* ctx, float* lse, float* headw, float* align, void* workspace,
                  int B, int T, int D, int H, void* stream);

/* Backward of the above w.r.t. x, query, att (closed form): one pass that reads
 * x and writes dx.  g_out [B,dh] f32 nullable (gradient of `out`);  g_ctx [B,H,dh] f32 nullable
 * (gradient arriving directly on ctx, i.e. MultiHeadAttention used alone);  ctx/lse/headw are
 * the forward's saved outputs (headw nullable iff att is NULL);  dx [B,T,D] in x_dtype;
 * dquery [dh,H] f32, datt [dh] f32 (nullable iff att NULL) are OVERWRITTEN, deterministically
 * (per-CTA partials in `workspace`, then a fixed-order reduction).
 */
size_t dasv_dmha_bwd_workspace_bytes(int B, int T, int D, int H);
int dasv_dmha_bwd(const void* x, int x_dtype, const int32_t* lengths,
                  const float* query, const float* att,
                  const float* g_out, const float* g_ctx,
                  const float* ctx, const float* lse, const float* headw,
                  void* dx, float* dquery, float* datt, void* workspace,
                  int B, int T, int D, int H, void* stream);

/* Attention.forward (scripts/poolings.py:22-27): single query over time, no scale; also the
 * stand-alone HeadAttention.forward (poolings.py:45-51,61-71; T = heads, D = head size) with its
 * training-mode drop mask.  x [B,T,D] (x_dtype), keep [B,T] uint8 nullable, att [D] f32,
 * out [B,D] f32, align [B,T] f32 (required: it is both the module's second return value and
 * the kernel's score workspace). */
int dasv_attention_fwd(const void* x, int x_dtype, const int32_t* lengths, const uint8_t* keep,
                       const float* att, float* out, float* align, int B, int T, int D, void* stream);

/* ---------------------------------------------------------------- VGG front-end
 * Activations are NHWC [B,T,F,C]; the reference's NCHW [B,1,T,80] input (CNNs.py:70) is NHWC
 * with C=1.  `lengths` (nullable) = valid frames per utterance AT THIS LAYER'S INPUT resolution;
 * rows t >= lengths[b] of the conv output are written as zero (SURVEY.md §5.7 masking rule).
 */

/* conv11: Conv2d(1, Cout, 3, padding 1) + bias + ReLU (CNNs.py:72).
 * x [B,T,F] f32, w [Cout,1,3,3] f32 (reference layout), bias [Cout] f32, y [B,T,F,Cout] (y_dtype). */
int dasv_conv11_direct(const float* x, const float* w, const float* bias, const int32_t* lengths,
                       void* y, int y_dtype, int B, int T, int F, int Cout, void* stream);
/* The same layer with LAZY masking (inference pipelines, see DASV_CONV_LAZY_MASK): rows t > lengths[b] of y may be left
 * unwritten; rows < lengths[b] and the zero row t = lengths[b] are as above. */
int dasv_conv11_direct_lazy(const float* x, const float* w, const float* bias, const int32_t* lengths,
                            void* y, int y_dtype, int B, int T, int F, int Cout, void* stream);

/* Weight re-packing (done once per module, cached by the caller):
 *   f32 path : w [Cout,Cin,3,3] f32 -> [9][Cin][Cout] f32
 *   bf16 path: w [Cout,Cin,3,3] f32 -> [Cout_pad][9][Cin] bf16 (K-major A operand for tcgen05;
 *              Cout_pad = Cout rounded up to 128, extra rows zero; element count from
 *              dasv_packed_conv_weight_bf16_elems) */
int dasv_pack_conv_weight_f32(const float* w, float* packed, int Cout, int Cin, void* stream);
size_t dasv_packed_conv_weight_bf16_elems(int Cout, int Cin);
int dasv_pack_conv_weight_bf16(const float* w, void* packed, int Cout, int Cin, void* stream);
/* The same packing with the 16-bit format chosen: dtype DASV_BF16 or DASV_F16 (three more mantissa bits; weights sit far
 * inside fp16's range).  Pass DASV_CONV_W_F16 to the convolution for fp16-packed weights. */
int dasv_pack_conv_weight_16(const float* w, void* packed, int Cout, int Cin, int dtype, void* stream);
/* fp32-parity mode on the tensor cores ("fp32x3"): an fp32 value v travels as two bf16 numbers hi = bf16(v),
 * lo = bf16(v - hi); w*x ~= hi(w)hi(x) + hi(w)lo(x) + lo(w)hi(x) with fp32 accumulation (error ~2^-16 per product).
 * Weights: [Cout_pad][9][3*Cin] bf16 = per tap [hi | hi | lo] (3x dasv_packed_conv_weight_bf16_elems elements).
 * Activations: NHWC with 2*C channels [hi(C) | lo(C)] (dasv_conv11_direct with y_dtype DASV_SPLIT_BF16, and the
 * NHWC output of dasv_conv3x3_igemm_bf16 with DASV_CONV_X3).  Three MMAs per product: a third of the bf16 rate. */
int dasv_pack_conv_weight_x3(const float* w, void* packed, int Cout, int Cin, void* stream);

/* fp32 CUDA-core implicit-GEMM conv3x3 + bias + ReLU (+ row mask): the fp32-parity path
 * (1e-4 relative) for conv12..conv42 (CNNs.py:73-85).  x [B,T,F,Cin] f32, wp packed f32,
 * y [B,T,F,Cout] f32.  Cin % 4 == 0, Cout % 4 == 0. */
int dasv_conv3x3_f32(const float* x, const float* wp, const float* bias, const int32_t* lengths,
                     float* y, int B, int T, int F, int Cin, int Cout, void* stream);

/* max_pool2d(2, stride 2, ceil_mode=True) on NHWC (CNNs.py:74,78,82,86); x [B,T,F,C] (dtype),
 * y [B,ceil(T/2),ceil(F/2),C] same dtype, or with ref_layout != 0 the front-end's final
 * [B,T',C*F'] tensor with feature index c*F'+f (CNNs.py:88-89) in y_dtype. */
int dasv_maxpool2x2(const void* x, int x_dtype, void* y, int y_dtype, int ref_layout,
                    int B, int T, int F, int C, void* stream);

/* bf16 tensor-core implicit-GEMM conv3x3 (tcgen05.mma, TMEM accumulators, TMA-fed) with bias,
 * ReLU, row mask and optionally the 2x2 ceil-mode max-pool fused into the epilogue
 * (CNNs.py:73-86, one call per conv).  x [B,T,F,Cin] bf16, wp [Cout][9][Cin] bf16, bias f32.
 * Output: without POOL y [B,T,F,Cout] bf16; with POOL y [B,T2,F2,Cout] bf16, or with
 * REF_LAYOUT y [B,T2,Cout*F2] in y_dtype (T2 = ceil(T/2), F2 = F/2).
 * Operand formats: the MMA is tcgen05 kind::f16, whose operands are both bf16 (default) or both fp16
 * (DASV_CONV_W_F16 | DASV_CONV_X_F16: wp, x and an NHWC y are fp16, y_dtype = DASV_F16); a mixed pair is rejected
 * (it is an illegal instruction on B200).  Accumulation is fp32 either way.
 * Requirements: Cin % 64 == 0, Cout % 8 == 0, F even and <= 256 (the reference's 80-bin input
 * gives F = 80, 40, 20, 10).  Needs the CUDA driver; the plan and both tensor maps are cached per
 * (x, wp, shape, flags), and the launch is a programmatic dependent of the stream's previous kernel. */
int dasv_conv3x3_igemm_bf16(const void* x, const void* wp, const float* bias, const int32_t* lengths,
                            void* y, int y_dtype, int flags,
                            int B, int T, int F, int Cin, int Cout, void* workspace, void* stream);
/* Small batches: a launch whose tiles would leave most SMs idle is split along K (every tile's K slices are dealt to
 * several CTAs, partial sums pass through `workspace`, a second streaming kernel adds them and applies bias / ReLU /
 * pool / format; deterministic).  Bytes of `workspace` needed for these arguments (0 = not split; NULL is then fine): */
size_t dasv_conv3x3_igemm_workspace_bytes(int y_dtype, int flags, int has_lengths, int B, int T, int F, int Cin, int Cout);

/* conv11 + conv12 of the front-end in ONE kernel (CNNs.py:72-74): x0 [B,T,F] f32, w11 [C1,1,3,3] f32, b11 [C1] f32, then
 * exactly dasv_conv3x3_igemm_bf16 on relu(conv11(x0) + b11) with wp [Cout][9][C1], bias, lengths, y, y_dtype, flags
 * (DASV_CONV_RELU | DASV_CONV_POOL | the operand formats).  The C1-channel tensor never exists in HBM: four extra warps
 * compute it per tile into a per-CTA scratch patch (L2-resident) one tile ahead of the MMAs.  Bit-identical to the two
 * separate calls.  C1 in {64, 128}; scratch: dasv_conv12_fused_workspace_bytes(C1) bytes, 128-byte aligned. */
size_t dasv_conv12_fused_workspace_bytes(int C1);
int dasv_conv12_fused_bf16(const float* x0, const float* w11, const float* b11, const void* wp, const float* bias,
                           const int32_t* lengths, void* y, void* scratch, int y_dtype, int flags,
                           int B, int T, int F, int C1, int Cout, void* stream);

/* Input gradient of the same convolution: dx = conv3x3(g, W') with the rotated, transposed weights
 * W'[ci][co][ky][kx] = W[co][ci][2-ky][2-kx] packed by dasv_pack_conv_weight_bf16 (wp_rot), linear epilogue, on the
 * forward's tensor-core kernel.  g [B,T,F,Cg] bf16 (Cg = the conv's output channels), dx [B,T,F,Cx] bf16.
 * relu_mask (nullable, bf16 [B,T,F,Cx]): the activation that fed this conv, i.e. the ReLU output of the layer below;
 * dx is zeroed where it is <= 0 (the ReLU backward of that layer fused into the store).  Rows t >= lengths[b] are 0. */
int dasv_conv3x3_dgrad_bf16(const void* g, const void* wp_rot, const void* relu_mask, const int32_t* lengths,
                            void* dx, int B, int T, int F, int Cg, int Cx, void* stream);

/* Weight gradient of a 3x3 stride-1 pad-1 convolution on the tensor cores (training side of the layer above; the
 * reference obtains it from autograd/cuDNN, scripts/CNNs.py:59-66):
 *   dw[co][ci][ky][kx] (+)= sum_{b,t,f} g[b,t,f,co] * x[b,t+ky-1,f+kx-1,ci]
 * x [B,T,F,Cin] bf16 (the layer's input), g [B,T,F,Cout] bf16 (gradient at the conv output, i.e. after the ReLU / pool
 * backward; zero for frames past an utterance), dw [Cout,Cin,3,3] f32 (reference layout; overwritten, or added to when
 * accumulate != 0); db [Cout] f32 nullable: the bias gradient sum_{b,t,f} g[b,t,f,co], from one extra MMA per K step
 * against a tile of ones.  Deterministic: split-K partials in `workspace` are added in fixed order.
 * Requirements: Cin % 64 == 0, Cout % 64 == 0, F <= 254. */
size_t dasv_conv3x3_wgrad_workspace_bytes(int B, int T, int F, int Cin, int Cout);
int dasv_conv3x3_wgrad_bf16(const void* x, const void* g, float* dw, float* db, void* workspace, int accumulate,
                            int B, int T, int F, int Cin, int Cout, void* stream);

/* Memory-bound pieces of the front-end's backward pass (what autograd does around the conv backward in the reference,
 * scripts/CNNs.py:72-86).  bf16 NHWC activations/gradients, f32 parameter gradients, deterministic sums.
 *  relu_bwd:        g [n] = y [n] > 0 ? g : 0, in place (n % 8 == 0).
 *  unpool_relu_bwd: backward of relu + max_pool2d(2,2,ceil_mode=True): y [B,T,F,C] = the conv's ReLU output before the
 *                   pool, gp = gradient at the pooled output (bf16 [B,T2,F2,C], or with gp_ref_layout_f32 the front-end's
 *                   f32 [B,T2,C*F2] output layout); writes all of g [B,T,F,C] (first maximum of a window gets gp if > 0).
 *  bias_grad:       db [C] (+)= column sums of g [P,C].
 *  conv11_bwd:      conv11 (Cin = 1): dw [C,1,3,3], db [C] from x [B,T,F] f32 and g [B,T,F,C]; input rows >= lengths[b]
 *                   count as zero, like the forward.
 * workspaces: dasv_bias_grad_workspace_bytes(C), dasv_conv11_bwd_workspace_bytes(B, T, C). */
int dasv_relu_bwd_bf16(void* g, const void* y, size_t n, void* stream);
int dasv_unpool_relu_bwd_bf16(const void* gp, int gp_ref_layout_f32, const void* y, void* g, int B, int T, int F, int C, void* stream);
size_t dasv_bias_grad_workspace_bytes(int C);
size_t dasv_conv11_bwd_workspace_bytes(int B, int T, int C);
int dasv_bias_grad_bf16(const void* g, float* db, void* workspace, int accumulate, size_t P, int C, void* stream);
int dasv_conv11_bwd(const float* x, const void* g, const int32_t* lengths, float* dw, float* db, void* workspace, int accumulate,
                    int B, int T, int F, int C, void* stream);

/* ---------------------------------------------------------------- embedding tail
 * getEmbedding's FC block, eval mode: b2(relu(fc2(relu(fc1(pooled))))) (scripts/model.py:56-57).
 * Packing: w1t [Din,E] = fc1.weight^T, w2t [E,E] = fc2.weight^T (f32),
 * bn_scale = b2.weight / sqrt(b2.running_var + eps), bn_shift = b2.bias - b2.running_mean*bn_scale.
 * pooled [B,Din] f32 -> emb [B,E] f32. */
int dasv_fc_tail_f32(const float* pooled, const float* w1t, const float* b1, const float* w2t,
                     const float* b2, const float* bn_scale, const float* bn_shift, float* emb,
                     int B, int Din, int E, void* stream);

/* ---------------------------------------------------------------- trial scoring
 * scoreCosineDistance = F.cosine_similarity(dim=-1, eps=1e-8) (scripts/utils.py:18-21), batched:
 *   pairs : score[i] = cos(emb[ia[i]], emb[ib[i]])            (the trial-list form, train.py:117-133)
 *   matrix: score[i,j] = cos(enrol[i], test[j])               (cross-product form)            */
int dasv_cosine_pairs(const float* emb, const int32_t* ia, const int32_t* ib, float* scores,
                      int n_pairs, int E, void* stream);
size_t dasv_cosine_matrix_workspace_bytes(int Ne, int Nt);
int dasv_cosine_matrix(const float* enrol, const float* test, float* scores, void* workspace,
                       int Ne, int Nt, int E, void* stream);

/* EER sweep counts (scripts/train.py:135-150 with scripts/utils.py:5-15): ge_counts[k] = #{scores[i] >= thresholds[k]},
 * compared in double like the reference; FAR = 100*ge/n on impostor scores, FRR = 100*(n-ge)/n on client scores.
 * scores [n] f32, thresholds [n_th] f64 (device), ge_counts [n_th] u64 (device, overwritten). */
int dasv_threshold_counts(const float* scores, int n, const double* thresholds, int n_th,
                          unsigned long long* ge_counts, void* stream);

/* ---------------------------------------------------------------- feature extraction (in front of the path)
 * scripts/featureExtractor.py:8-23 (`mfsc`): y*scale, pre-emphasis over the whole signal (first sample times 1 - preem),
 * frames of 512 samples (the reference's n_fft) every `hop`, `window` [win_length <= 512] centred in the frame
 * (librosa.stft(center=False)), |rfft|, mel = melw [n_mels,257] . |S| (librosa.filters.mel, built by the caller for its
 * sample rate; mel_range [n_mels,2] = first / one-past-last bin with a non-zero weight), log(max(1, mel)).
 * wave [B][wave_stride] f32, n_samples [B] int32; out [B,Tmax,n_mels] f32: frame t of utterance b is written iff
 * t < 1 + (n_samples[b] - 512) / hop, other rows are left untouched. */
int dasv_logmel_f32(const float* wave, const int32_t* n_samples, int B, long long wave_stride,
                    const float* window, int win_length, int hop,
                    const float* melw, const int32_t* mel_range, int n_mels,
                    float preem, float scale, float* out, int Tmax, void* stream);

/* scripts/featureExtractor.py:25-26 (`normalize`) / data.py:21-30 (`normalizeFeatures` 'cmn', and 'cmvn' with
 * variance != 0), in place: per utterance and mel bin subtract the mean over its frames[b] frames and, for 'cmvn', divide
 * by the population standard deviation where it exceeds 0.01 (data.py:28-29); rows t >= frames[b] are set to zero. */
int dasv_cmn_f32(float* feat, const int32_t* frames, int B, int Tmax, int n_mels, int variance, void* stream);

/* ---------------------------------------------------------------- training-mode tail (scripts/model.py:61-71)
 * BatchNorm1d with batch statistics (model.py:67): y = (x - mean) * rsqrt(var_biased + eps) * gamma + beta over the batch
 * dimension of x [B,E]; running_mean / running_var (nullable) are updated in place like torch.nn.BatchNorm1d
 * (momentum, unbiased variance).  save_mean / save_invstd [E] feed the backward. */
int dasv_bn1d_train_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                        float* y, float* save_mean, float* save_invstd, int B, int E, float eps, float momentum, void* stream);
int dasv_bn1d_train_bwd(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_invstd,
                        float* dx, float* dgamma, float* dbeta, int B, int E, void* stream);

/* AM-Softmax (scripts/loss.py:37-52): costh = normalise(x) @ normalise_columns(W), logits = s * costh - margin_scaled at
 * [b, label[b]] (margin_scaled = s * m / (1 + alpha), alpha = the annealing term of loss.py:28-35).  x [B,E], W [E,S],
 * label [B] int64 ON THE DEVICE (the reference scatters the margin on the CPU, loss.py:45-48).  inv_x [B] / inv_w [S]
 * (inverse norms) are outputs kept for the backward, which takes the gradients at both outputs (either nullable). */
int dasv_amsoftmax_fwd(const float* x, const float* W, const long long* label, float* costh, float* logits,
                       float* inv_x, float* inv_w, int B, int E, int S, float s, float margin_scaled, void* stream);
size_t dasv_amsoftmax_bwd_workspace_bytes(int B, int S);
int dasv_amsoftmax_bwd(const float* dcosth, const float* dlogits, const float* x, const float* W, const float* costh,
                       const float* inv_x, const float* inv_w, float* dx, float* dW, void* workspace,
                       int B, int E, int S, float s, void* stream);

/* Debugging aid: the conv3x3_igemm launches that follow write eight %globaltimer stamps (ns) per CTA into
 * buf[blockIdx.x * 8 + i] (0 CTA start, 1 set-up done, 2 previous kernel of the stream finished, 3 first operands landed,
 * 4 last MMA issued, 5 first accumulator complete, 6 epilogue done, 7 CTA end); buf = device memory for 8 * 148 * 2 values,
 * NULL (the default) switches the stamps off.  Process-wide, not for concurrent use. */
int dasv_debug_conv_trace(unsigned long long* buf);

/* ---------------------------------------------------------------- host-side helper of the variable-length extractor
 * (the reference embeds one utterance per call, scripts/train.py:117-133; the extractor batches them): n host->device
 * copies on `stream`, one cudaMemcpyAsync each -- segment i = nbytes[i] bytes from (char*)src_host + src_off[i] to
 * (char*)dst + dst_off[i].  src_host should be pinned (the copies then overlap the kernels of other streams); the three
 * offset/size arrays are HOST arrays, read before the call returns. */
int dasv_h2d_segments(void* dst, const void* src_host, const long long* src_off, const long long* dst_off,
                      const long long* nbytes, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DASV_B200_H */
