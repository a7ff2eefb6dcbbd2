/*
 * mqgan_b200.h — C ABI of libmqgan_b200.so, the sm_100a (B200) kernels behind the
 * PreEncoder re-encode path of ZDisket/MQGAN.
 *
 * The reference has no FFI: every op on this path is a PyTorch/ATen call inside
 * preencoder.py / attentions.py / quantizer.py.  Each entry point below names the
 * reference call site(s) (file:line under the reference root) whose arithmetic it
 * replaces.  All pointers are DEVICE pointers unless marked host; the caller owns
 * every buffer (inputs, outputs, workspace) and passes the CUDA stream to launch
 * on.  Functions return 0 on success, non-zero on error (1 = bad argument,
 * 2 = CUDA error, 3 = unsupported device); mq_last_error() returns a
 * thread-local message.  Nothing throws across this boundary and the library
 * keeps no per-call global state (re-entrant across per-GPU workers).
 *
 * Layouts are channel-last everywhere: 1-D activations (B, T, C); refiner
 * activations (B, T', F, C).  "row_mask" arrays are uint8, one byte per (b, t)
 * row at that tensor's time resolution, 1 = padded (the reference's x_mask,
 * preencoder.py:15-24).
 */
#ifndef MQGAN_B200_H_
#define MQGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mq_stream_t; /* cudaStream_t */

#define MQ_MAX_TAPS 16
#define MQ_MAX_SEGS 6

/* ---- library ------------------------------------------------------------- */
int mq_version(void);                 /* 10000*major + 100*minor + patch */
const char* mq_last_error(void);      /* thread-local, never NULL */
/* 0 if the current device is sm_100 (B200); 3 otherwise. */
int mq_device_check(void);
int mq_sm_count(void);

/* ---- K1/K3/K4/K9/K10/K11: implicit-GEMM convolution on tcgen05 ------------ */
/*
 * out[n,h,w,co] = epi( sum_tap sum_seg sum_c  in[n, h+dh[tap], w+dw[tap], a_coff[seg]+c]
 *                                           * wpack[co, ((seg*taps+tap)*kchunks*64 + c)] )
 * Zero padding comes from TMA out-of-bounds fill.  bf16 operands, fp32 accumulate
 * in TMEM.  nseg = 1 is a plain bf16 convolution; nseg = 6 with the input stored
 * as three bf16 terms [x0|x1|x2] along channels is the fp32-grade "bf16x3" mode
 * (products x0w0,x0w1,x1w0,x0w2,x1w1,x2w0).
 *
 * Replaces: F.linear preencoder.py:433,486,490; F.conv1d attentions.py:532,533,541
 * (same padding) and :471-474 (causal); F.conv2d preencoder.py:97-98; with the
 * following element-wise tail fused: bias; masked_fill attentions.py:536-537,547-548,
 * preencoder.py:101; APTx attentions.py:34-35; residual adds attentions.py:545,
 * preencoder.py:99-100.
 *
 * epilogue:  v = acc + bias[co]
 *            if res_mode == 1: v += res           (ResidualBlock1D: before mask/act)
 *            if mask_pre  && row_mask[n*H+h]: v = 0
 *            if act: v = (1 + tanh(beta v)) * gamma * v
 *            if res_mode == 2: v += res           (ConvBlock: y + x after act)
 *            if mask_post && row_mask[n*H+h]: v = 0
 */
typedef struct mq_conv_params {
  /* input activation: bf16, (N, H, W, in_ld) channel-last, in_ld = channel pitch */
  const void* in;
  int N, H, W;
  int in_ld;
  /* packed weights: bf16 [cout_pad][K], K = nseg*taps*kchunks*64 (segment-major), K-major */
  const void* wpack;
  int cout;      /* real output channels */
  int cout_pad;  /* rows of wpack: multiple of bn */
  int bn;        /* N tile: multiple of 32, <= 256 */
  int taps, nseg, kchunks;
  int tap_dh[MQ_MAX_TAPS];
  int tap_dw[MQ_MAX_TAPS];
  int a_coff[MQ_MAX_SEGS];
  int bh, bw;    /* pixel sub-tile, bh*bw <= 128 */
  int msub;      /* sub-tiles (stacked along H) per CTA tile sharing one weight tile: 1, 2 or 4
                    with msub*bn <= 256; 0 = 1.  Cuts L2->SM weight traffic for narrow layers. */
  /* epilogue */
  const float* bias;       /* [cout] or NULL */
  const uint8_t* row_mask; /* [N*H] or NULL */
  int mask_pre, mask_post;
  int act;                 /* 0 none, 1 APTx */
  int fast_tanh;           /* 1: tanh.approx, 0: ~1e-7 abs-error tanh */
  float beta, gamma;
  int res_mode;            /* 0 none, 1 before mask/act, 2 after act */
  const void* res;         /* (N,H,W,res_ld) */
  int res_is_bf16, res_ld, res_coff;
  float* out_f32;  int f32_ld, f32_coff;      /* optional fp32 output */
  void* out_bf16;  int bf16_ld, bf16_coff;    /* optional bf16 output */
  void* out_split; int split_ld, split_seg;   /* optional bf16x3 output: term j at channel j*split_seg + co */
  /* Fused nearest Upsample((2,1)) + channel concat of UpBlock (preencoder.py:123-130), optional.
   * in2 != NULL: `in` is the LOW-resolution tensor (N, H, W, in_ld), in2 the skip tensor
   * (N, 2H, W, in2_ld); the output (and row_mask / res / out_*) has 2H rows.  Output row 2i+p
   * reads up(x)[2i+p+dh] = x[(2i+p+dh)>>1], so per row parity p the three row taps collapse to two
   * with pre-summed weights: taps [0, up_taps) read `in` at half-row offset tap_dh (p = 0) or
   * tap_dh_odd (p = 1) with kchunks chunks each; taps [up_taps, taps) read in2 at output-row
   * offset tap_dh with kchunks2 chunks each.  wpack holds [2][cout_pad][K] (parity-major),
   * K = (up_taps*kchunks + (taps-up_taps)*kchunks2)*64.  nseg must be 1. */
  const void* in2;
  int in2_ld, up_taps, kchunks2;
  int tap_dh_odd[MQ_MAX_TAPS];
  /* f16x2 mode: op_f16 != 0 means `in` / `in2` / `wpack` hold fp16 (not bf16) operands;
   * split_kind selects what out_split writes: 0 = three bf16 terms, 1 = two fp16 terms
   * (term j at channel j*split_seg + co).  With nseg = 3 segments h1*g0, h0*g1, h0*g0 this is the
   * cheaper fp32-grade mode (22-bit operands, half the MMA work of bf16x3). */
  int op_f16, split_kind;
  /* acc_scale: the accumulator is multiplied by this before the bias is added (0 means 1).
   * f16x2 weights are packed pre-scaled by a power of two so their low terms stay in fp16's
   * normal range; acc_scale = 2^-s undoes it exactly. */
  float acc_scale;
  /* halo != 0: halo-tile main loop for single-source 3x3 / pad-1 convolutions (taps in row-major
   * (dh, dw) order, nseg == 1, bh == 16, bw == 8): each 64-channel chunk of the activation is
   * fetched once per CTA tile as a (16*msub + 2) x 10 pixel halo and the nine taps are nine
   * shifted tensor-core descriptors into it (9x less L2->SM activation traffic). */
  int halo;
  /* pair != 0: CTA-pair main loop (tcgen05 cta_group::2, M = 256) for single-source 3x3 / pad-1
   * convolutions and for the fused upsample-concat mode (bh == 16, bw == 8, nseg == 1): the two
   * CTAs of a cluster each stage the halo of their own msub sub-tiles and half of every weight
   * tile, halving weight traffic per pixel (L2->SM and shared-memory reads).  Excludes halo.
   * With W == 1 (1-D convolutions / linears; bh == 128, bw == 1, msub == 1, consecutive row taps, any
   * nseg) the row-halo variant runs: one (128 + taps - 1)-row activation fetch per channel chunk, taps
   * as descriptors shifted by whole rows. */
  int pair;
  /* out_pool != NULL: the epilogue also writes AvgPool2d((2,1)) of the bf16 output, filled with zero
   * where either source row is padded (preencoder.py:111-114 with the max-pooled mask :63-65, :96):
   * (N, H/2, W, pool_ld) bf16, channel c at c (no offset).  Needs the bf16-only epilogue (out_bf16,
   * cout % 32 == 0, no fp32/split output), bw == 8, bh even, H even; not with in2. */
  void* out_pool;
  int pool_ld;
} mq_conv_params;

int mq_conv_gemm(const mq_conv_params* p, mq_stream_t stream);

/* ---- input staging: fp32 -> bf16 / bf16x3 (feeds K1) ---------------------- */
/* x (rows, C) fp32 -> out (rows, nterms*C) 16-bit terms, term j at [j*C, (j+1)*C):
 * nterms 1 = bf16, 3 = bf16x3, 2 = f16x2 (two fp16 terms). */
int mq_split_bf16(const float* x, void* out, int64_t rows, int C, int nterms, mq_stream_t stream);

/* ---- K2: ConvBlock2D `pre` / `post` (preencoder.py:277-301) ---------------- */
/*
 * x (B, T, C) -> y (B, T, C):  s = dw5x5 over the (channel, time) plane (+bias),
 * zero at padded rows; y = sum_k wout[k] * aptx(wpw[k]*s + bpw[k]; 1, .5) + bout,
 * and y = bout at padded rows.  The (B, C, C, T) expansion of the reference
 * (preencoder.py:288-295) is never materialised.
 * dw: 25 floats [i over channel][j over time] then the dw bias (26 total, folded
 * weight-norm); pw: C x {wpw, bpw, wout} interleaved as float4-padded triples.
 *
 * y depends on the pixel only through the scalar s, so the C-term sum
 * g(s) = sum_k wout[k]*aptx(wpw[k]*s + bpw[k]) + bout is a fixed univariate function
 * of the model.  If `table` is non-NULL it holds per-interval cubic coefficients of g
 * on a uniform grid: interval i covers s*table_inv_h in [i - table_off, i - table_off + 1),
 * g ~= c0 + t*(c1 + t*(c2 + t*c3)), t = frac(s*table_inv_h); table_inv_h must be a power
 * of two so t is exact.  Pixels whose s falls outside the grid use the exact C-term sum.
 * The host builds the table in float64 and verifies it against the exact sum
 * (mqgan_b200/engine.py:_CB2D); with table == NULL every pixel is evaluated exactly.
 */
typedef struct mq_cb2d_params {
  const void* x; int x_is_bf16;   /* fp32 or bf16 input */
  int B, T, C;
  const float* dw;                /* [26] */
  const float* pw;                /* [C][4]: wpw, bpw, wout, 0 */
  float bout;
  const uint8_t* row_mask;        /* [B*T] or NULL */
  int fast_tanh;
  float* out_f32;                 /* optional (B,T,C) */
  void* out_bf16;                 /* optional (B,T,C) */
  void* out_split;                /* optional (B,T,3C) bf16x3 */
  const float* table;             /* optional [table_n][4] cubic coefficients of g */
  int table_n, table_off;
  float table_inv_h;
  int split_kind;                 /* out_split format: 0 = bf16x3 (B,T,3C), 1 = f16x2 (B,T,2C) */
} mq_cb2d_params;
int mq_convblock2d(const mq_cb2d_params* p, mq_stream_t stream);

/* ---- K5/K6: CBAM1D (attentions.py:217-273, 310-365, 393-419) + block tail --- */
/* pass 1: per (b, c) max over ALL t (quirk App. B1) and masked sum over valid t,
 * deterministic two-stage reduction.  o (B,T,C) fp32; part: workspace
 * [B][nchunk][2][C] fp32 with nchunk = mq_cam_chunks(T). */
int mq_cam_chunks(int T);
int mq_cam_reduce(const float* o, const uint8_t* row_mask, int B, int T, int C, float* part,
                  mq_stream_t stream);
/* pass 1b: finish the reduction, shared MLP on max and mean, sigmoid -> gate (B,C).
 * w0 (R,C), b0 (R), w2 (C,R), b2 (C). */
int mq_cam_gate(const float* part, const uint8_t* row_mask, int B, int T, int C, int R,
                const float* w0, const float* b0, const float* w2, const float* b2, float* gate,
                mq_stream_t stream);
/* pass 2: SAM pools (max / mean over C of gate*o, unmasked), conv k7 (no bias),
 * sigmoid; y = aptx(mask(sam*gate*o + o + r); beta, gamma)   (attentions.py:545-549).
 * r: residual (B,T,C) fp32.  Outputs: y fp32 and/or bf16 / bf16x3. */
typedef struct mq_cbam_apply_params {
  const float* o; const float* gate; const float* res;
  const uint8_t* row_mask;
  int B, T, C;
  const float* sam_w;  /* [2][7] */
  float beta, gamma;
  float* out_f32; void* out_bf16; void* out_split;
  int split_kind;      /* 0 = bf16x3, 1 = f16x2 */
} mq_cbam_apply_params;
int mq_cbam_apply(const mq_cbam_apply_params* p, mq_stream_t stream);

/* ---- K7: q_in_proj + FSQ (preencoder.py:448-451, quantizer.py:109-114,137,177-181) */
/* y (rows, C) fp32 -> z = y W^T + b (fp32 FMA), bound/round-half-even/mixed-radix
 * index -> idx int64 (rows).  Optionally also writes z (rows, D) fp32.  D <= 8.
 * half_l / shift / offset / half_w / basis: host arrays of D entries. */
typedef struct mq_fsq_params {
  int D;
  float half_l[8], shift[8], offset[8];
  int half_w[8], basis[8], levels[8];
} mq_fsq_params;
int mq_qin_fsq(const float* y, int64_t rows, int C, const float* w, const float* b,
               const mq_fsq_params* fsq, int64_t* idx, float* z_out, mq_stream_t stream);
/* FSQ alone on latents z (rows, D): the quantizer.py eval path. */
int mq_fsq_quantize(const float* z, int64_t rows, const mq_fsq_params* fsq, int64_t* idx,
                    float* codes_out, mq_stream_t stream);

/* ---- nearest-codeword lookup: distance GEMM + argmin + gather (BASELINE configs[3]) ---------- */
/*
 * idx[i] = argmin_k ||z_i - c_k||^2 (ties -> lowest k), codes_out[i] = c_idx[i], dist_out[i] =
 * ||c||^2 - 2 z.c at the minimum (i.e. the squared distance minus ||z_i||^2).  The (n x k) distance
 * matrix is never materialised: z.c^T is a tcgen05 GEMM (128 latents x 256 codes per accumulator),
 * the ||c||^2 bias, the argmin and the gather run in its epilogue.
 *
 * The reference's quantiser is FSQ (quantizer.py:109-181, restated exactly by mq_fsq_quantize); it
 * has no learned codebook.  Its implicit codebook (quantizer.py:101-104) is one valid `codebook`
 * here, and the two agree except on exact ties; any other (k, d <= 64) fp32 codebook works too.
 *
 * cb_img / c2 / acc_scale come from the one-time host-side packing (mqgan_b200/ops.py:pack_codebook):
 *   cb_img  [k_pad / (256*slices)][nterm][256][128 B] ready-made 128-byte-swizzled shared-memory
 *           images of the codebook operand, slices = 4 / ks codes side by side per row,
 *           ks = K-steps of 16 per code = 1, 2 or 4 (>= ceil(d / 16)); nterm = 1 (bf16) or 2 (fp16 h0, h1);
 *   c2      [k_pad] fp32 ||c_k||^2, +inf on the padding codes;
 *   acc_scale = 1 / (power-of-two pre-scale applied to the f16x2 codebook terms), 0 means 1.
 * mode 0: bf16 operands, one product (index agreement with fp32 is reported, not exact);
 * mode 1: "f16x2", three fp16 products of 2-term splits (22-bit operands): fp32-grade argmin.
 */
typedef struct mq_vq_params {
  const float* z;          /* (n, d) fp32 latents, read once */
  int64_t n;
  int d;                   /* 1..64 */
  const void* cb_img;
  const float* c2;
  const float* codebook;   /* (k, d) fp32, gathered into codes_out */
  int k, k_pad;
  int mode;
  float acc_scale;
  int64_t* idx;            /* (n) */
  float* codes_out;        /* (n, d) or NULL */
  float* dist_out;         /* (n) or NULL */
  /* fold != 0: the packed codebook carries -2 c (times the pre-scale) and, in three extra K columns d .. d+2 of every
   * code, the terms of ||c||^2 * pre-scale / zconst; the kernel writes zconst into the same columns of the latent
   * operand, so the accumulator is the score itself and the epilogue is a bare running minimum (ops.py:pack_codebook
   * chooses this whenever d + 3 fits the code's K-steps, i.e. d <= 13, 29 or 61).  c2 is then unused. */
  int fold;
  float zconst;
} mq_vq_params;
int mq_vq_nearest(const mq_vq_params* p, mq_stream_t stream);

/* Measurement aid for mq_vq_nearest's roofline (bench.py secondary.vq_lookup): one CTA per SM, eight warps that do
 * nothing but read their 128 lanes x 512 columns of tensor memory with tcgen05.ld (the epilogue's access pattern,
 * no MMA, no score math) `iters` times.  Bytes read = sm_count * iters * 128 * 512 * 4; the caller times the launch
 * and gets the chip's TMEM read bandwidth - the ceiling of any kernel that must read every one of the n x k fp32
 * scores from TMEM once.  `sink` (1 float, device) only keeps the loads alive.  No reference counterpart. */
int mq_tmem_read_probe(int iters, float* sink, mq_stream_t stream);

/* ---- K8: indices_to_codes + q_out_proj as a table gather (quantizer.py:183-205,
 *          preencoder.py:464-466) -------------------------------------------- */
/* table (n_codes, C) fp32 = q_out_proj(implicit_codebook); idx (rows) int64;
 * out (rows, C) bf16 and/or fp32.  Indices outside [0, n_codes) are an error flag
 * written to *bad (device int, optional) and the row is zero-filled. */
int mq_code_gather(const int64_t* idx, int64_t rows, const float* table, int n_codes, int C,
                   void* out_bf16, float* out_f32, int* bad, mq_stream_t stream);

/* ---- a19/K13: refiner masks and resampling ---------------------------------- */
/* mask (B,T) -> all refiner masks.  Level l has H_l = T8 >> l rows per batch
 * element, T8 = T rounded up to a multiple of 2^depth.  Both outputs are flat
 * uint8 buffers holding levels 0..depth back to back; level l starts at byte
 * offset B * sum_{j<l} H_j  (total B * sum_l H_l bytes each).
 *   down[0]   = mask padded with ones to T8             (preencoder.py:29-47)
 *   down[l]   = max-pool(2) of down[l-1]                (preencoder.py:63-65)
 *   up[depth] = down[depth];  up[l] = nearest-up(2) of up[l+1]   (preencoder.py:68-70)
 * mask may be NULL (= all valid, preencoder.py:471-472). */
int mq_refiner_masks(const uint8_t* mask, int B, int T, int depth, uint8_t* down, uint8_t* up,
                     mq_stream_t stream);
/* Zero every row (row_bytes each, multiple of 16) with mask_new[row] == 1 and mask_old[row] == 0
 * (mask_old may be NULL).  Re-masks a skip tensor from the down-path mask to the coarser up-path
 * mask (preencoder.py:125, :96) before the fused upsample+concat convolution reads it. */
int mq_zero_rows(void* x, const uint8_t* mask_new, const uint8_t* mask_old, int64_t rows,
                 int64_t row_bytes, mq_stream_t stream);
/* AvgPool2d((2,1)) + masked_fill by the pooled mask (preencoder.py:111-114, :96).
 * x (B, H, F, C) bf16 -> y (B, H/2, F, C) bf16; mask_out (B*H/2). */
int mq_avgpool_mask(const void* x, void* y, const uint8_t* mask_out, int B, int H, int F, int C,
                    mq_stream_t stream);
/* Upsample((2,1), nearest) + cat([up, skip], channel) + masked_fill (preencoder.py:123-130, :96).
 * x (B, H/2, F, Cx), skip (B, H, F, Cs) -> y (B, H, F, Cx+Cs) bf16. */
int mq_upcat_mask(const void* x, const void* skip, void* y, const uint8_t* mask_out, int B, int H,
                  int F, int Cx, int Cs, mq_stream_t stream);
/* The same two passes for the fp32-grade decoder mode (decoder_precision "f16x2"), whose activations are two fp16
 * terms [h0 | h1] along channels (2C per pixel, x = h0 + h1 to 22 bits).  The pool averages h0 + h1 of both rows in
 * fp32 - the reference pools in fp32, preencoder.py:112 - and re-splits; the concat is a pure copy into
 * [x_h0 | skip_h0 | x_h1 | skip_h1], the term layout mq_conv_gemm's a_coff expects with term stride Cx + Cs. */
int mq_avgpool_mask_split(const void* x, void* y, const uint8_t* mask_out, int B, int H, int F, int C,
                          mq_stream_t stream);
int mq_upcat_mask_split(const void* x, const void* skip, void* y, const uint8_t* mask_out, int B, int H,
                        int F, int Cx, int Cs, mq_stream_t stream);

/* ---- K12: thin refiner convolutions --------------------------------------- */
/* refiner.pre.conv1 (1 -> C, 3x3, pad 1) + APTx on the masked, T-padded refiner
 * input (preencoder.py:172-175, 96-97).  r (B, T, F) fp32 = cat[x_recon, hidden];
 * w (C, 9) folded, b (C).  y (B, T8, F, C) bf16. */
int mq_refiner_stem(const float* r, const uint8_t* mask, int B, int T, int T8, int F, int C,
                    const float* w, const float* b, int fast_tanh, void* y, mq_stream_t stream);
/* The same with precise tanh and a two-term fp16 output y (B, T8, F, 2C) = [h0 | h1] (fp32-grade decoder mode). */
int mq_refiner_stem_split(const float* r, const uint8_t* mask, int B, int T, int T8, int F, int C,
                          const float* w, const float* b, void* y, mq_stream_t stream);
/* refiner.post (C -> 1, 3x3) + crop + mask + reproj (F -> M, no bias) + x_recon add
 * (preencoder.py:191-200, 499).  The C -> 1 3x3 convolution is split into a 1x1 tcgen05 GEMM
 * C -> 9 (mq_conv_gemm with the (9, C) folded weight, tap k = 3*(dt+1) + (df+1)) producing
 * taps (B, T8, F, ldp) fp32, and this kernel: post[t,f] = bias + sum_k taps[t+dt, f+df, k]
 * (zero outside the image), crop to T, mask, reproj_t (F, M) = reproj.weight transposed,
 * out (B, T, M) fp32 = r[..., :M] + residual, r (B, T, F) = cat[x_recon, hidden].
 * ldp == 4: taps (B, T8, F, 4) holds three ROW sums per pixel, channel dt + 1 = sum over df and c of kernel row dt
 * applied at that pixel's own row (mq_conv_gemm, "taps2d" weight (4, C, 3) with taps dh = 0, dw = -1, 0, 1), and
 * post[t,f] = bias + taps[t-1,f,0] + taps[t,f,1] + taps[t+1,f,2]: what the engine uses.
 * ldp == 1: taps (B, T8, F, 1) already holds the 3x3 sum (mq_conv_gemm with the (1, C, 3, 3) weight packed as
 * "conv2d3", one output channel) and post[t,f] = bias + taps[t, f]. */
int mq_refiner_tail(const float* taps, int ldp, const uint8_t* mask, int B, int T, int T8, int F,
                    float bias, const float* reproj_t, int M, const float* r, float* out,
                    mq_stream_t stream);

/* ---- f3: log-mel front-end (convert_spectrograms.py:14-35) ------------------ */
/*
 * out[b, f, m] = log(max(sum_k fb[k, m] * |STFT(wav_b)[f, k]|, clip)): torchaudio
 * MelSpectrogram(n_fft, win_length, hop_length, f_min, f_max, n_mels, power=1) with its defaults
 * (hann window, center=True, pad_mode="reflect", onesided, htk mel scale, norm=None) followed by
 * log(clamp(min=1e-5)), laid out (frames, n_mels) as convert_spectrograms.py:35 returns it.
 * wav (B, wav_ld) fp32 with lengths[b] valid samples; utterance b has 1 + lengths[b] / hop frames
 * (0 if lengths[b] <= n_fft/2, which torch.stft rejects); rows beyond that up to out_frames are zero.
 * Host-prepared tables (mqgan_b200/melspec.py): window [n_fft] (hann(win_length) centred in n_fft),
 * twiddle [n_fft][2] = (cos, -sin)(2 pi t / n_fft), and the mel filterbank in sparse row form:
 * mel bin m sums fb_w[fb_off[m] + i] * |S[fb_start[m] + i]| for i < fb_count[m].
 */
typedef struct mq_melspec_params {
  const float* wav;
  int64_t wav_ld;
  const int64_t* lengths;   /* device, [B] */
  int B;
  int n_fft, hop, n_mels, n_freqs;
  const float* window;
  const float* twiddle;
  const int* fb_start;
  const int* fb_count;
  const int* fb_off;
  const float* fb_w;
  float clip;
  float* out;               /* (B, out_frames, n_mels) fp32 */
  int64_t out_frames;
} mq_melspec_params;
int mq_log_mel(const mq_melspec_params* p, mq_stream_t stream);

/* ---- f4 (training step): weight gradient of the convolutions on tcgen05 ------ */
/*
 * dw[s][tap][co][ci] = sum over the pixels of K-split s of  dy[n,h,w,co] * x[n, h+tap_dh[tap], w+tap_dw[tap], ci]
 * (zero outside the image): the weight gradient autograd produces for F.conv2d preencoder.py:97-98,
 * F.conv1d attentions.py:471-474, 532-541 and F.linear preencoder.py:433,486,490 inside
 * train.py:380-501 (loss.backward()).  dy (N,H,W,dy_ld) and x (N,H,W,x_ld) are bf16 channel-last; the
 * contraction over pixels runs on tcgen05 with both operands MN-major (no transposes), fp32 accumulate.
 * The caller sums the `split` partial tensors (deterministic; no atomics).  bh*bw must be 64: the pixel
 * box of one K block (8x8 for images, 64x1 for sequences).  mq_conv_wgrad_split() returns the split the
 * kernel wants for a problem (fills 148 SMs); any split >= 1 is valid.
 * The data gradient needs no kernel of its own: it is mq_conv_gemm on dy with the taps mirrored and the
 * weight transposed (mqgan_b200/training.py:dgrad_pack).
 */
typedef struct mq_wgrad_params {
  const void* dy; int dy_ld;
  const void* x;  int x_ld;
  int N, H, W;
  int cout, cin;
  int taps;
  int tap_dh[MQ_MAX_TAPS];
  int tap_dw[MQ_MAX_TAPS];
  int bh, bw;
  int split;
  float* dw;       /* [split][taps][cout][cin] fp32 */
} mq_wgrad_params;
int mq_conv_wgrad_split(const mq_wgrad_params* p);
int mq_conv_wgrad(const mq_wgrad_params* p, mq_stream_t stream);

/* ---- f4 (training step): ConvBlock2D point-wise stage, forward and backward ---- */
/*
 * y[p] = bout + sum_k wout[k] * aptx(wpw[k]*s[p] + bpw[k]; 1, .5) at valid rows, bout at padded rows
 * (preencoder.py:288-295 after the masked depth-wise conv :286-287), for s (rows, C) fp32 with one
 * row_mask byte per row; parameters are DEVICE arrays [C] (bout: [1]) because they change every step.
 * mq_cb2d_backward returns ds (rows, C) = dL/ds and per-block partial sums part[blocks][3][C] of
 * (dL/dwpw, dL/dbpw, dL/dwout), blocks = mq_cb2d_grad_blocks(rows, C); dL/dbout = sum(dy) is the caller's.
 * fast_tanh != 0: tanh.approx.f32 (one MUFU op, ~2^-11 relative) instead of the ~1e-7 ex2 + rcp form; the kernels
 * are MUFU-bound (C tanh per pixel).  The (B, C, C, T) expansion autograd would keep alive in the reference (train.py:380-501 -> backward of
 * preencoder.py:288-295) never exists.
 */
int mq_cb2d_point_forward(const float* s, const uint8_t* row_mask, int64_t rows, int C, const float* wpw,
                          const float* bpw, const float* wout, const float* bout, int fast_tanh, float* y,
                          mq_stream_t stream);
int mq_cb2d_grad_blocks(int64_t rows, int C);
int mq_cb2d_backward(const float* s, const float* dy, const uint8_t* row_mask, int64_t rows, int C,
                     const float* wpw, const float* bpw, const float* wout, int fast_tanh, float* ds, float* part,
                     mq_stream_t stream);

/* ---- f4 (training step): fused activation passes of the refiner's ConvBlock ----- */
/*
 * u (pixels, C) fp32 convolution output; bf16 everywhere else; row_mask[pixel / pix_per_row] != 0 = padded row.
 * forward:  out = padded ? 0 : (1 + tanh(beta u)) gamma u [+ res]        (preencoder.py:97-101, attentions.py:34-35)
 * backward: du = padded ? 0 : dy * d aptx/du (u),  dres (optional) = padded ? 0 : dy,  dbias_part (optional) below
 * i.e. what autograd runs for APTx -> (+x) -> masked_fill in train.py:400/484's backward.  C % 8 == 0.
 */
int mq_act_forward(const float* u, const void* res_bf16, const uint8_t* row_mask, int64_t pixels, int C,
                   int pix_per_row, float beta, float gamma, void* out_bf16, mq_stream_t stream);
int mq_act_backward(const void* dy_bf16, const float* u, const uint8_t* row_mask, int64_t pixels, int C,
                    int pix_per_row, float beta, float gamma, void* du_bf16, void* dres_bf16, float* dbias_part,
                    mq_stream_t stream);
/* dbias_part (optional): [mq_act_bias_blocks(pixels, C)][C] per-block column sums of du in fp32 - summed over blocks they
 * are the bias gradient of the convolution that produced u.  mq_act_bias_blocks returns 0 when C does not allow it. */
int mq_act_bias_blocks(int64_t pixels, int C);

/* Discriminator activation (discriminators.py:234, 247): out = pix_mask[pixel] ? 0 : LeakyReLU_slope(u + bias) over a
 * channels-last (pixels, C) tensor, u fp32 or bf16 (cuDNN's autocast output, computed WITHOUT its bias), bias fp32 [C] or
 * NULL, out / dy / du bf16; one pass each way.  dbias_part (optional): [mq_act_bias_blocks(pixels, C)][C] per-block column
 * sums of du = the conv bias gradient, so neither cuDNN nor a reduction kernel has to re-read du for it. */
int mq_leaky_mask_forward(const void* u, int u_is_bf16, const float* bias, const uint8_t* pix_mask, int64_t pixels, int C,
                          float slope, void* out_bf16, mq_stream_t stream);
int mq_leaky_mask_backward(const void* dy_bf16, const void* u, int u_is_bf16, const float* bias, const uint8_t* pix_mask,
                           int64_t pixels, int C, float slope, void* du_bf16, float* dbias_part, mq_stream_t stream);

/* ---- f2: .npy file I/O of the re-encode CLI (host code; reencode_spectrograms.py:49-62, 69-81) ------------ */
/*
 * Foreign calls run without the Python GIL, so the CLI's I/O threads scale; np.load / np.save do not for small files.
 * mq_npy_probe: shape of a C-order 2-D little-endian float32 / float64 / float16 .npy (dtype 0 / 1 / 2).
 * mq_npy_read_f32: reads min(rows, dst_rows) rows into dst (dst_rows, cols) as float32 and zero-fills the rest (the
 * reference's zero padding to the batch's longest utterance); *rows_out = rows in the file.
 * mq_npy_write_f32: writes src (rows, cols) float32 exactly as np.save would (format 1.0, 64-byte aligned data).
 * Return 0 ok, 1 bad argument / column mismatch, 4 unsupported file (the caller falls back to numpy), 5 I/O error.
 */
int mq_npy_probe(const char* path, int64_t* rows, int64_t* cols, int* dtype, int64_t* data_offset);
int mq_npy_read_f32(const char* path, float* dst, int64_t dst_rows, int64_t cols, int64_t* rows_out);
int mq_npy_write_f32(const char* path, const float* src, int64_t rows, int64_t cols);

/* ---- sequence mask (preencoder.py:15-24) ----------------------------------- */
int mq_sequence_mask(const int64_t* lengths, int B, int T, uint8_t* mask, mq_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MQGAN_B200_H_ */
