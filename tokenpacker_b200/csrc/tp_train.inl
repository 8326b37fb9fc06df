// Training path of the projector: forward that keeps what the backward needs, and the backward itself.
// Included by tp_api.cu inside its extern "C" block (shares the launch helpers of that translation unit).
//
// The reference trains this module (it is the only trainable module of stage 1 and is trained in stage 2:
// llava/train/train.py:950-958, scripts/v1_5/pretrain*.sh), through PyTorch autograd over builder.py:107-137.  Here:
//   forward_train = the inference launch plan without the out_proj fold, and with the pre-activations z of the two GELUs kept in
//                   memory (GELU'(z) needs z; GELU(z) is not invertible): the epilogues of k/v_proj.0 and mlp.0 store z and GELU(z) together;
//   backward      = dgrad GEMMs in the NN form of the pair kernel (dY . W, the weight read as stored), wgrad GEMMs in the TN form
//                   (dW = dY^T . X straight from the row-major activations: MN-major UMMA tiles, contraction over rows),
//                   LayerNorm / GELU / window-attention backward kernels, bias gradients as deterministic column sums.
// Gradients w.r.t. the CLIP features are not produced (the tower is frozen in every released recipe; the Python layer raises
// if the inputs require grad).

}  // extern "C"  (reopened below)

namespace {

struct SavedLayout {
  size_t z_kv, h_kv;        // [R,2048]
  size_t y_k, y_v;          // [R,1024]
  size_t stats;             // f32 [(2R+Q), 8, 2]
  size_t k_p, v_p;          // [R,1024]
  size_t q, y_q, q_p, ctx, o;   // [Q,1024]
  size_t z_m, h_m;          // [Q,H]
  size_t total;
};

SavedLayout saved_layout(long long n_crops, int s, int H) {
  const size_t R = static_cast<size_t>(n_crops) * kTokens;
  const int g = kGrid / s;
  const size_t Q = static_cast<size_t>(n_crops) * g * g;
  SavedLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  L.z_kv = take(R * 2 * kC * 2); L.h_kv = take(R * 2 * kC * 2);
  L.y_k = take(R * kC * 2); L.y_v = take(R * kC * 2);
  L.stats = take((2 * R + Q) * kStatSlots * 2 * 4);
  L.k_p = take(R * kC * 2); L.v_p = take(R * kC * 2);
  L.q = take(Q * kC * 2); L.y_q = take(Q * kC * 2); L.q_p = take(Q * kC * 2); L.ctx = take(Q * kC * 2); L.o = take(Q * kC * 2);
  L.z_m = take(Q * static_cast<size_t>(H) * 2); L.h_m = take(Q * static_cast<size_t>(H) * 2);
  L.total = off;
  return L;
}

struct BwdLayout {
  size_t w_m2t, w_m0t, w_ot, w_iqt, w_ikt, w_ivt, w_k2t, w_v2t;   // transposed weight (dgrad fallback: only w_m2t is ever allocated)
  size_t g_t, hm_t, dzm, dzm_t, o_t, d_o, do_t, ctx_t, dctx;
  size_t dqp, dkp, dvp, dqp_t, dkp_t, dvp_t, lnq_t, lnk_t, lnv_t;
  size_t dqh, dkh, dvh, dyq, dyk, dyv, dyq_t, dyk_t, dyv_t, q_t, hkv_t;
  size_t dzkv, dzkv_t, xm_t;
  size_t ln_part;        // f32 [3][kLnBlocks][3][1024]
  size_t col_part;       // f32 [kColChunks][max(H, 2048)]  column-sum partials (bias gradients)
  size_t splitk;         // f32 [2][kWgradSplits][1024,1024]  split-K partial sums of the 1024x1024 wgrads that contract over R rows
  size_t total;
  long long Rp, Qp;
};

constexpr int kLnBlocks = 296;
constexpr int kColChunks = 592;     // 4 CTAs per SM: each sums ~rows/592 rows of 1024 columns
constexpr int kWgradSplits = 4;     // 16 output tiles x 4 K-slices = 64 tiles of R/4 rows instead of 16 tiles of R rows

BwdLayout bwd_layout(long long n_crops, int s, int H) {
  const size_t R = static_cast<size_t>(n_crops) * kTokens;
  const int g = kGrid / s;
  const size_t Q = static_cast<size_t>(n_crops) * g * g;
  BwdLayout L;
  L.Rp = static_cast<long long>(align_up(R, 8));
  L.Qp = static_cast<long long>(align_up(Q, 8));
  const size_t Qp = L.Qp, Hs = H;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  // transposed weight copies: only the dgrad fallback needs one (n_in % 256 != 0 = mlp.2 at tiny hidden sizes); every other dgrad reads
  // the weight as stored (NN form)
  L.w_m2t = take((H % 256 != 0) ? Hs * Hs * 2 : 0);
  L.w_m0t = L.w_ot = L.w_iqt = L.w_ikt = L.w_ivt = L.w_k2t = L.w_v2t = 0;
  // scratch for the transposing wgrad fallback: only mlp.2's weight gradient with hidden % 256 != 0 ever takes it
  const size_t fb = (H % 256 != 0) ? Hs * Qp * 2 : 0;
  L.g_t = take(fb); L.hm_t = take(fb);
  L.dzm = take(Q * Hs * 2);
  L.d_o = take(Q * kC * 2); L.dctx = take(Q * kC * 2);
  L.dqp = take(Q * kC * 2); L.dkp = take(R * kC * 2); L.dvp = take(R * kC * 2);
  L.lnq_t = take(Q * kC * 2); L.lnk_t = take(R * kC * 2); L.lnv_t = take(R * kC * 2);      // LayerNorm outputs (not transposed)
  L.dqh = take(Q * kC * 2); L.dkh = take(R * kC * 2); L.dvh = take(R * kC * 2);
  L.dyq = take(Q * kC * 2); L.dyk = take(R * kC * 2); L.dyv = take(R * kC * 2);
  L.dzkv = take(R * 2 * kC * 2);
  L.dzm_t = L.o_t = L.do_t = L.ctx_t = L.dqp_t = L.dkp_t = L.dvp_t = L.dyq_t = L.dyk_t = L.dyv_t = L.q_t = L.hkv_t = L.dzkv_t = L.xm_t = 0;   // unused (TN wgrad)
  L.ln_part = take(3ull * kLnBlocks * 3 * kC * 4);
  L.col_part = take(static_cast<size_t>(kColChunks) * (Hs > 2048 ? Hs : 2048) * 4);
  L.splitk = take(2ull * kWgradSplits * kC * kC * 4);
  L.total = off;
  return L;
}

int launch_transpose(const void* in, long long ld_in, void* out, long long ld_out, long long rows, int cols, cudaStream_t stream) {
  const dim3 grid(static_cast<unsigned>((cols + 31) / 32), static_cast<unsigned>((rows + 31) / 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, stream>>>(static_cast<const __nv_bfloat16*>(in), ld_in, static_cast<__nv_bfloat16*>(out), ld_out, rows,
                                                      cols);
  TP_CUDA(cudaGetLastError()); ++g_launch_count;
  return TP_OK;
}

int launch_gelu_fwd(const void* z, void* h, size_t elems, cudaStream_t stream) {
  const long long n8 = static_cast<long long>(elems / 8);
  gelu_fwd_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(z),
                                                                                static_cast<__nv_bfloat16*>(h), n8);
  TP_CUDA(cudaGetLastError()); ++g_launch_count;
  return TP_OK;
}

int launch_ln_bwd(const void* g, const void* y, const float* stats, const void* gamma, void* dy, float* partial, long long rows,
                  void* dgamma, void* dbeta, void* dbias, cudaStream_t stream) {
  ln_bwd_kernel<<<kLnBlocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(g), static_cast<const __nv_bfloat16*>(y), stats,
                                               static_cast<const __nv_bfloat16*>(gamma), static_cast<__nv_bfloat16*>(dy), partial, rows);
  TP_CUDA(cudaGetLastError()); ++g_launch_count;
  ln_param_reduce_kernel<<<kC / 32, dim3(32, kReduceLanes), 0, stream>>>(partial, kLnBlocks, static_cast<__nv_bfloat16*>(dgamma),
                                                       static_cast<__nv_bfloat16*>(dbeta), static_cast<__nv_bfloat16*>(dbias));
  TP_CUDA(cudaGetLastError()); ++g_launch_count;
  return TP_OK;
}

GemmItem plain_item(const void* a, long long lda, const void* b, long long ldb, void* c, long long ldc, long long M, long long N, long long K,
                    const float* bias = nullptr, float alpha = 1.0f) {
  GemmItem it{AOperand{a, lda, 0, 0}, b, ldb, M, N, K, plain_epilogue(c, ldc, bias, 0)};
  it.ep.alpha = alpha;
  return it;
}

}  // namespace

extern "C" {

size_t tp_train_saved_bytes(int64_t n_crops, int scale_factor, int hidden) {
  if (n_crops <= 0 || scale_factor <= 0 || kGrid % scale_factor != 0 || !valid_hidden(hidden)) return 0;
  return saved_layout(n_crops, scale_factor, hidden).total;
}

size_t tp_backward_workspace_bytes(int64_t n_crops, int scale_factor, int hidden) {
  if (n_crops <= 0 || scale_factor <= 0 || kGrid % scale_factor != 0 || !valid_hidden(hidden)) return 0;
  return bwd_layout(n_crops, scale_factor, hidden).total;
}

int tp_forward_train(const tp_weights* w, const void* packed, const void* x0, const void* xm, int64_t n_crops, int64_t x0_crop_stride, int64_t xm_crop_stride,
                     int scale_factor, int hidden, void* out, void* saved, size_t saved_bytes, void* stream_) {
  if (scale_factor <= 0 || kGrid % scale_factor != 0) return TP_ERR_BAD_SCALE_FACTOR;
  if (packed == nullptr || x0 == nullptr || xm == nullptr || out == nullptr || saved == nullptr || n_crops <= 0 || !valid_hidden(hidden))
    return TP_ERR_INVALID_ARGUMENT;
  if (x0_crop_stride < static_cast<int64_t>(kTokens) * kC || xm_crop_stride < static_cast<int64_t>(kTokens) * kCm || x0_crop_stride % 8 != 0 ||
      xm_crop_stride % 8 != 0 || n_crops * kTokens > 0x7fff0000ll)
    return TP_ERR_INVALID_ARGUMENT;
  DeviceInfo dev;
  TP_TRY(device_info(&dev));
  const int s = scale_factor, H = hidden;
  const int g = kGrid / s, Mq = g * g;
  const long long R = n_crops * kTokens, Q = n_crops * Mq;
  const SavedLayout S = saved_layout(n_crops, s, H);
  if (saved_bytes < S.total) return TP_ERR_WORKSPACE_TOO_SMALL;
  const PackedLayout L = packed_layout(H);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const uint8_t* P = static_cast<const uint8_t*>(packed);
  uint8_t* sv = static_cast<uint8_t*>(saved);
  auto wf = [&](size_t off) { return reinterpret_cast<const float*>(P + off); };
  auto bf = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(sv + off); };
  // weight matrices that need no transformation: the live parameters in place (w given), else their copies in the packed buffer
  const void* W_k2 = w != nullptr ? w->k_proj_2_w : P + L.w_k2;
  const void* W_v2 = w != nullptr ? w->v_proj_2_w : P + L.w_v2;
  const void* W_q = w != nullptr ? w->q_proj_w : P + L.w_q;
  const void* W_o = w != nullptr ? w->out_proj_w : P + L.w_o;
  const void* W_m0 = w != nullptr ? w->mlp_0_w : P + L.w_m0;
  const void* W_m2 = w != nullptr ? w->mlp_2_w : P + L.w_m2;
  if (w != nullptr && (W_k2 == nullptr || W_v2 == nullptr || W_q == nullptr || W_o == nullptr || W_m0 == nullptr || W_m2 == nullptr))
    return TP_ERR_INVALID_ARGUMENT;
  float* stats_k = reinterpret_cast<float*>(sv + S.stats);
  float* stats_v = stats_k + 2 * kStatSlots * R;
  float* stats_q = stats_v + 2 * kStatSlots * R;

  // GELU stages: bit 0 = k/v_proj.0, bit 1 = mlp.0 store z and GELU(z) from one epilogue (else a GEMM that stores z + an elementwise
  // GELU pass).  TP_TRAIN_DUAL overrides (A/B aid, read per call); the dual store needs the two-slab staging of the default build.
  int dual_mask = (Gemm2Config::kOutBufs == 2 && kSlabCols == 64 && dev.sms >= 2) ? 3 : 0;
  if (const char* e = getenv("TP_TRAIN_DUAL")) dual_mask = (Gemm2Config::kOutBufs == 2 && kSlabCols == 64 && dev.sms >= 2) ? atoi(e) : 0;
  {
    const __nv_bfloat16* x0p = static_cast<const __nv_bfloat16*>(x0);
    TP_TRY(launch_front_s(s, x0p, x0_crop_stride, bf(S.q), Q, stream));
  }
  {
    AOperand a{xm, kCm, 0, 0};
    if (xm_crop_stride != static_cast<int64_t>(kTokens) * kCm) a = AOperand{xm, kCm, kTokens, xm_crop_stride};
    if (dual_mask & 1) {
      // one pass: the epilogue stores the rounded pre-activation z (kept for GELU'(z)) AND GELU(z)
      GemmItem it{a, P + L.w_kv0, kCm, R, 2 * kC, kCm, plain_epilogue(bf(S.h_kv), 2 * kC, wf(L.b_kv0), 1)};
      it.ep.dual = 1;
      it.c_pre = bf(S.z_kv);
      it.ld_pre = 2 * kC;
      TP_TRY(launch_gemms(&it, 1, dev.sms, stream));
    } else {
      TP_TRY(launch_gemm(a, P + L.w_kv0, kCm, R, 2 * kC, kCm, plain_epilogue(bf(S.z_kv), 2 * kC, wf(L.b_kv0), 0), dev.sms, stream));
      TP_TRY(launch_gelu_fwd(bf(S.z_kv), bf(S.h_kv), static_cast<size_t>(R) * 2 * kC, stream));
    }
  }
  {
    GemmItem gi[3];
    gi[0] = GemmItem{AOperand{bf(S.h_kv), 2 * kC, 0, 0}, W_k2, kC, R, kC, kC, plain_epilogue(bf(S.y_k), kC, wf(L.b_k2), 0)};
    gi[0].ep.stats_out = stats_k; gi[0].ep.stats_out_slots = kStatSlots;
    gi[1] = GemmItem{AOperand{bf(S.h_kv) + kC, 2 * kC, 0, 0}, W_v2, kC, R, kC, kC, plain_epilogue(bf(S.y_v), kC, wf(L.b_v2), 0)};
    gi[1].ep.stats_out = stats_v; gi[1].ep.stats_out_slots = kStatSlots;
    gi[2] = GemmItem{AOperand{bf(S.q), kC, 0, 0}, W_q, kC, Q, kC, kC, plain_epilogue(bf(S.y_q), kC, nullptr, 0)};
    gi[2].ep.stats_out = stats_q; gi[2].ep.stats_out_slots = kStatSlots;
    TP_TRY(launch_gemms(gi, 3, dev.sms, stream));
  }
  {
    GemmItem gi[3];
    gi[0] = GemmItem{AOperand{bf(S.y_k), kC, 0, 0}, P + L.w_ik, kC, R, kC, kC, plain_epilogue(bf(S.k_p), kC, wf(L.c_k), 0)};
    gi[0].ep.col_a = wf(L.wsum_k); gi[0].ep.stats_in = stats_k; gi[0].ep.stats_in_slots = kStatSlots;
    gi[1] = GemmItem{AOperand{bf(S.y_v), kC, 0, 0}, P + L.w_iv, kC, R, kC, kC, plain_epilogue(bf(S.v_p), kC, wf(L.c_v), 0)};
    gi[1].ep.col_a = wf(L.wsum_v); gi[1].ep.stats_in = stats_v; gi[1].ep.stats_in_slots = kStatSlots;
    gi[2] = GemmItem{AOperand{bf(S.y_q), kC, 0, 0}, P + L.w_iq, kC, Q, kC, kC, plain_epilogue(bf(S.q_p), kC, wf(L.c_q), 0)};
    gi[2].ep.col_a = wf(L.wsum_q); gi[2].ep.stats_in = stats_q; gi[2].ep.stats_in_slots = kStatSlots;
    gi[2].ep.alpha = 0.08838834764831845f;
    TP_TRY(launch_gemms(gi, 3, dev.sms, stream));
  }
  TP_TRY(launch_attn_s(s, bf(S.q_p), bf(S.k_p), bf(S.v_p), bf(S.ctx), Q, stream));
  TP_TRY(launch_gemm(AOperand{bf(S.ctx), kC, 0, 0}, W_o, kC, Q, kC, kC, plain_epilogue(bf(S.o), kC, wf(L.b_o), 0), dev.sms, stream));
  if ((dual_mask & 2) && H % 256 == 0) {
    GemmItem it{AOperand{bf(S.o), kC, 0, 0}, W_m0, kC, Q, H, kC, plain_epilogue(bf(S.h_m), H, wf(L.b_m0), 1)};
    it.ep.dual = 1;
    it.c_pre = bf(S.z_m);
    it.ld_pre = H;
    TP_TRY(launch_gemms(&it, 1, dev.sms, stream));
  } else {
    TP_TRY(launch_gemm(AOperand{bf(S.o), kC, 0, 0}, W_m0, kC, Q, H, kC, plain_epilogue(bf(S.z_m), H, wf(L.b_m0), 0), dev.sms, stream));
    TP_TRY(launch_gelu_fwd(bf(S.z_m), bf(S.h_m), static_cast<size_t>(Q) * H, stream));
  }
  TP_TRY(launch_gemm(AOperand{bf(S.h_m), H, 0, 0}, W_m2, H, Q, H, H, plain_epilogue(out, H, wf(L.b_m2), 0), dev.sms, stream));
  return TP_OK;
}

int tp_backward(const tp_weights* w, const void* xm, int64_t xm_crop_stride, int64_t n_crops, int scale_factor, int hidden,
                const void* grad_out, const void* saved, const tp_weights* grads, void* workspace, size_t workspace_bytes, void* stream_) {
  if (w == nullptr || grads == nullptr || xm == nullptr || grad_out == nullptr || saved == nullptr || workspace == nullptr || n_crops <= 0 ||
      !valid_hidden(hidden))
    return TP_ERR_INVALID_ARGUMENT;
  if (scale_factor <= 0 || kGrid % scale_factor != 0) return TP_ERR_BAD_SCALE_FACTOR;
  if (xm_crop_stride != static_cast<int64_t>(kTokens) * kCm) return TP_ERR_INVALID_ARGUMENT;   // backward takes contiguous xm
  {
    const void* const* f = reinterpret_cast<const void* const*>(w);
    const void* const* gptr = reinterpret_cast<const void* const*>(grads);
    for (size_t i = 0; i < sizeof(tp_weights) / sizeof(void*); ++i)
      if (f[i] == nullptr || gptr[i] == nullptr) return TP_ERR_INVALID_ARGUMENT;
  }
  DeviceInfo dev;
  TP_TRY(device_info(&dev));
  const int s = scale_factor, H = hidden;
  const int g = kGrid / s, Mq = g * g;
  const long long R = n_crops * kTokens, Q = n_crops * Mq;
  const SavedLayout S = saved_layout(n_crops, s, H);
  const BwdLayout B = bwd_layout(n_crops, s, H);
  if (workspace_bytes < B.total) return TP_ERR_WORKSPACE_TOO_SMALL;
  const long long Rp = B.Rp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const uint8_t* sv = static_cast<const uint8_t*>(saved);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto sb = [&](size_t off) { return reinterpret_cast<const __nv_bfloat16*>(sv + off); };
  auto wb = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  const float* stats_k = reinterpret_cast<const float*>(sv + S.stats);
  const float* stats_v = stats_k + 2 * kStatSlots * R;
  const float* stats_q = stats_v + 2 * kStatSlots * R;
  float* ln_part = reinterpret_cast<float*>(ws + B.ln_part);
  const __nv_bfloat16* in_w = static_cast<const __nv_bfloat16*>(w->in_proj_w);
  __nv_bfloat16* d_in_w = static_cast<__nv_bfloat16*>(const_cast<void*>(grads->in_proj_w));
  __nv_bfloat16* d_in_b = static_cast<__nv_bfloat16*>(const_cast<void*>(grads->in_proj_b));
  auto G = [&](const void* p) { return const_cast<void*>(p); };
  const float alpha_q = 0.08838834764831845f;

  // dgrad  dX[rows, n_in] = alpha * dY[rows, n_out] . W[n_out, n_in]: the NN form of the pair kernel reads the weight as stored (MN-major
  // B tiles).  Fallback for n_in not a multiple of 256 (mlp.2 at tiny hidden sizes): a transposed copy + the ordinary NT form.
  auto dgrad = [&](const void* dy, long long ld_dy, const void* wt, long long ld_w, void* dx, long long ld_dx, long long rows, int n_in,
                   int n_out, float alpha, size_t wt_off, GemmItem* item) -> int {
    if (n_in % 256 == 0) {
      *item = plain_item(dy, ld_dy, wt, ld_w, dx, ld_dx, rows, n_in, n_out, nullptr, alpha);
      item->tn = 2;
      return TP_OK;
    }
    TP_TRY(launch_transpose(wt, ld_w, wb(wt_off), n_out, n_out, n_in, stream));
    *item = plain_item(dy, ld_dy, wb(wt_off), n_out, dx, ld_dx, rows, n_in, n_out, nullptr, alpha);
    return TP_OK;
  };

  // wgrad  dW[n_out, n_in] = alpha * dY^T X  (dY: [rows, n_out], X: [rows, n_in], both as stored).  Normal case: the TN form of
  // the pair kernel reads both operands in place (MN-major UMMA tiles).  Fallback for n_in not a multiple of 256 (tiny hidden
  // sizes): explicit transposes into scratch + the NT kernel.
  auto wgrad = [&](const void* dy, long long ld_dy, const void* x, long long ld_x, long long rows, int n_out, int n_in, void* dw,
                   long long ld_dw, float alpha, size_t dy_t_off, size_t x_t_off, GemmItem* item) -> int {
    if (n_in % 256 == 0) {
      *item = GemmItem{AOperand{dy, ld_dy, 0, 0}, x, ld_x, n_out, n_in, rows, plain_epilogue(dw, ld_dw, nullptr, 0)};
      item->ep.alpha = alpha;
      item->tn = 1;
      return TP_OK;
    }
    const long long ldt = static_cast<long long>(align_up(static_cast<size_t>(rows), 8));
    TP_TRY(launch_transpose(dy, ld_dy, wb(dy_t_off), ldt, rows, n_out, stream));
    TP_TRY(launch_transpose(x, ld_x, wb(x_t_off), ldt, rows, n_in, stream));
    *item = plain_item(wb(dy_t_off), ldt, wb(x_t_off), ldt, dw, ld_dw, n_out, n_in, rows, nullptr, alpha);
    return TP_OK;
  };
  // The 1024x1024 wgrads that contract over all R = 576 N rows are 16 output tiles of 576 N / 64 k-blocks each: alone they leave
  // 58 of 74 CTA pairs idle for the longest GEMM time of the step.  They run split-K (kWgradSplits fp32 partial slices, summed in
  // fixed order by splitk_reduce_kernel) INSIDE the launch of the dgrads of the same stage, whose ~1300 tiles fill the machine.
  float* splitk = reinterpret_cast<float*>(ws + B.splitk);
  auto wgrad_split = [&](const void* dy, long long ld_dy, const void* x, long long ld_x, long long rows, int slot, GemmItem* item) -> bool {
    if (rows < 64ll * kBlockK * kWgradSplits) return false;            // short contractions: not worth the partials
    *item = GemmItem{AOperand{dy, ld_dy, 0, 0}, x, ld_x, kC, kC, rows, plain_epilogue(splitk + static_cast<size_t>(slot) * kWgradSplits * kC * kC, kC, nullptr, 0)};
    item->tn = 1;
    item->k_splits = kWgradSplits;
    item->ep.out_f32 = 1;
    return true;
  };
  auto wgrad_reduce = [&](int slot, float alpha, void* dw) -> int {
    const long long elems = static_cast<long long>(kC) * kC;
    splitk_reduce_kernel<<<static_cast<unsigned>((elems / 4 + 255) / 256), 256, 0, stream>>>(
        splitk + static_cast<size_t>(slot) * kWgradSplits * kC * kC, kWgradSplits, elems, alpha, static_cast<__nv_bfloat16*>(dw));
    TP_CUDA(cudaGetLastError()); ++g_launch_count;
    return TP_OK;
  };
  // bias gradient = column sums of dY (deterministic two-stage reduction: kColChunks row chunks, then a fixed-order sum)
  float* col_part = reinterpret_cast<float*>(ws + B.col_part);
  auto bias_grad = [&](const void* dy, long long ld_dy, long long rows, int cols, float scale, void* out) -> int {
    if (cols > (H > 2048 ? H : 2048) || cols % 8 != 0) return TP_ERR_INVALID_ARGUMENT;
    const int chunks = static_cast<int>(rows < kColChunks ? rows : kColChunks);
    colsum_partial_kernel<<<dim3((cols + 1023) / 1024, chunks), 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(dy), ld_dy, rows, cols,
                                                                                  chunks, col_part);
    TP_CUDA(cudaGetLastError()); ++g_launch_count;
    colsum_reduce_kernel<<<(cols + 31) / 32, dim3(32, kReduceLanes), 0, stream>>>(col_part, chunks, cols, cols, scale, static_cast<__nv_bfloat16*>(out));
    TP_CUDA(cudaGetLastError()); ++g_launch_count;
    return TP_OK;
  };
  // dz = dh GELU'(z) in place + bias gradient(s) = column sums of dz; out_hi != nullptr: columns [cols/2, cols) go to out_hi
  auto gelu_bwd_bias = [&](__nv_bfloat16* dh, const __nv_bfloat16* z, long long ld, long long rows, int cols, void* out, void* out_hi) -> int {
    if (cols > (H > 2048 ? H : 2048) || cols % 8 != 0) return TP_ERR_INVALID_ARGUMENT;
    const int chunks = static_cast<int>(rows < kColChunks ? rows : kColChunks);
    gelu_bwd_colsum_kernel<<<dim3((cols + 1023) / 1024, chunks), 128, 0, stream>>>(dh, z, ld, rows, cols, chunks, col_part);
    TP_CUDA(cudaGetLastError()); ++g_launch_count;
    const int n_out = out_hi != nullptr ? cols / 2 : cols;
    colsum_reduce_kernel<<<(n_out + 31) / 32, dim3(32, kReduceLanes), 0, stream>>>(col_part, chunks, n_out, cols, 1.0f, static_cast<__nv_bfloat16*>(out));
    TP_CUDA(cudaGetLastError()); ++g_launch_count;
    if (out_hi != nullptr) {
      colsum_reduce_kernel<<<(n_out + 31) / 32, dim3(32, kReduceLanes), 0, stream>>>(col_part + n_out, chunks, n_out, cols, 1.0f,
                                                                                      static_cast<__nv_bfloat16*>(out_hi));
      TP_CUDA(cudaGetLastError()); ++g_launch_count;
    }
    return TP_OK;
  };
  auto ln_apply = [&](const __nv_bfloat16* y, const float* stats, const void* gamma, const void* beta, __nv_bfloat16* out, long long rows) -> int {
    ln_apply_kernel<<<static_cast<unsigned>((rows * 32 + 255) / 256), 256, 0, stream>>>(y, stats, static_cast<const __nv_bfloat16*>(gamma),
                                                                                        static_cast<const __nv_bfloat16*>(beta), out, rows);
    TP_CUDA(cudaGetLastError()); ++g_launch_count;
    return TP_OK;
  };

  // ---- mlp.2:  out = h_m W_m2^T + b
  TP_TRY(bias_grad(grad_out, H, Q, H, 1.0f, G(grads->mlp_2_b)));
  {
    GemmItem gi[2];
    TP_TRY(wgrad(grad_out, H, sb(S.h_m), H, Q, H, H, G(grads->mlp_2_w), H, 1.0f, B.g_t, B.hm_t, &gi[0]));      // dW_m2 = G^T h_m
    TP_TRY(dgrad(grad_out, H, w->mlp_2_w, H, wb(B.dzm), H, Q, H, H, 1.0f, B.w_m2t, &gi[1]));                    // dh_m = G W_m2
    TP_TRY(launch_gemms(gi, 2, dev.sms, stream));
  }
  // ---- mlp.0:  z_m = o W_m0^T + b   (dz_m = dh_m GELU'(z_m) and the column sums of dz_m in one pass)
  TP_TRY(gelu_bwd_bias(wb(B.dzm), sb(S.z_m), H, Q, H, G(grads->mlp_0_b), nullptr));
  {
    GemmItem gi[2];
    TP_TRY(wgrad(wb(B.dzm), H, sb(S.o), kC, Q, H, kC, G(grads->mlp_0_w), kC, 1.0f, B.dzm_t, B.o_t, &gi[0]));   // dW_m0 = dz_m^T o
    TP_TRY(dgrad(wb(B.dzm), H, w->mlp_0_w, kC, wb(B.d_o), kC, Q, kC, H, 1.0f, B.w_m0t, &gi[1]));               // do = dz_m W_m0
    TP_TRY(launch_gemms(gi, 2, dev.sms, stream));
  }
  // ---- out_proj:  o = ctx W_o^T + b
  TP_TRY(bias_grad(wb(B.d_o), kC, Q, kC, 1.0f, G(grads->out_proj_b)));
  {
    GemmItem gi[2];
    TP_TRY(wgrad(wb(B.d_o), kC, sb(S.ctx), kC, Q, kC, kC, G(grads->out_proj_w), kC, 1.0f, B.do_t, B.ctx_t, &gi[0]));   // dW_o = do^T ctx
    TP_TRY(dgrad(wb(B.d_o), kC, w->out_proj_w, kC, wb(B.dctx), kC, Q, kC, kC, 1.0f, B.w_ot, &gi[1]));           // dctx = do W_o
    TP_TRY(launch_gemms(gi, 2, dev.sms, stream));
  }
  // ---- window attention
  {
    const long long threads = Q * 32;
    const unsigned blocks = static_cast<unsigned>((threads + 255) / 256);
    if (s == 2) window_attn_bwd_kernel<2><<<blocks, 256, 0, stream>>>(sb(S.q_p), sb(S.k_p), sb(S.v_p), wb(B.dctx), wb(B.dqp), wb(B.dkp), wb(B.dvp), Q);
    else if (s == 3) window_attn_bwd_kernel<3><<<blocks, 256, 0, stream>>>(sb(S.q_p), sb(S.k_p), sb(S.v_p), wb(B.dctx), wb(B.dqp), wb(B.dkp), wb(B.dvp), Q);
    else if (s == 4) window_attn_bwd_kernel<4><<<blocks, 256, 0, stream>>>(sb(S.q_p), sb(S.k_p), sb(S.v_p), wb(B.dctx), wb(B.dqp), wb(B.dkp), wb(B.dvp), Q);
    else window_attn_bwd_stream_kernel<<<blocks, 256, 0, stream>>>(sb(S.q_p), sb(S.k_p), sb(S.v_p), wb(B.dctx), wb(B.dqp), wb(B.dkp), wb(B.dvp), Q, s);
    TP_CUDA(cudaGetLastError()); ++g_launch_count;
  }
  // ---- MHA in-projections:  q' = alpha (LN(y_q) W_iq^T + b),  k' = LN(y_k) W_ik^T + b,  v' likewise
  TP_TRY(ln_apply(sb(S.y_q), stats_q, w->ln_q_w, w->ln_q_b, wb(B.lnq_t), Q));      // LN outputs, [rows,1024] as stored (not transposed)
  TP_TRY(ln_apply(sb(S.y_k), stats_k, w->ln_k_w, w->ln_k_b, wb(B.lnk_t), R));
  TP_TRY(ln_apply(sb(S.y_v), stats_v, w->ln_v_w, w->ln_v_b, wb(B.lnv_t), R));
  TP_TRY(bias_grad(wb(B.dqp), kC, Q, kC, alpha_q, d_in_b));
  TP_TRY(bias_grad(wb(B.dkp), kC, R, kC, 1.0f, d_in_b + kC));
  TP_TRY(bias_grad(wb(B.dvp), kC, R, kC, 1.0f, d_in_b + 2 * kC));
  {
    // one launch: the two long wgrads (split-K, first: longest tiles start earliest), the three dgrads, the short q wgrad
    GemmItem gi[6];
    int n = 0;
    const bool sk = wgrad_split(wb(B.dkp), kC, wb(B.lnk_t), kC, R, 0, &gi[0]) && wgrad_split(wb(B.dvp), kC, wb(B.lnv_t), kC, R, 1, &gi[1]);
    if (sk) {
      n = 2;
    } else {
      TP_TRY(wgrad(wb(B.dkp), kC, wb(B.lnk_t), kC, R, kC, kC, d_in_w + static_cast<size_t>(kC) * kC, kC, 1.0f, B.dkp_t, B.dyk_t, &gi[n++]));
      TP_TRY(wgrad(wb(B.dvp), kC, wb(B.lnv_t), kC, R, kC, kC, d_in_w + 2 * static_cast<size_t>(kC) * kC, kC, 1.0f, B.dvp_t, B.dyv_t, &gi[n++]));
    }
    TP_TRY(dgrad(wb(B.dkp), kC, in_w + static_cast<size_t>(kC) * kC, kC, wb(B.dkh), kC, R, kC, kC, 1.0f, B.w_ikt, &gi[n++]));      // d LN(y_k)
    TP_TRY(dgrad(wb(B.dvp), kC, in_w + 2 * static_cast<size_t>(kC) * kC, kC, wb(B.dvh), kC, R, kC, kC, 1.0f, B.w_ivt, &gi[n++]));
    TP_TRY(dgrad(wb(B.dqp), kC, in_w, kC, wb(B.dqh), kC, Q, kC, kC, alpha_q, B.w_iqt, &gi[n++]));                                 // d LN(y_q)
    TP_TRY(wgrad(wb(B.dqp), kC, wb(B.lnq_t), kC, Q, kC, kC, d_in_w, kC, alpha_q, B.dqp_t, B.q_t, &gi[n++]));
    TP_TRY(launch_gemms(gi, n, dev.sms, stream));
    if (sk) {
      TP_TRY(wgrad_reduce(0, 1.0f, d_in_w + static_cast<size_t>(kC) * kC));
      TP_TRY(wgrad_reduce(1, 1.0f, d_in_w + 2 * static_cast<size_t>(kC) * kC));
    }
  }
  // ---- LayerNorms
  // (the column sums of dy_k / dy_v = the bias gradients of k_proj_1.2 / v_proj_1.2 come out of the same pass; q_proj_1 has no bias)
  TP_TRY(launch_ln_bwd(wb(B.dqh), sb(S.y_q), stats_q, w->ln_q_w, wb(B.dyq), ln_part, Q, G(grads->ln_q_w), G(grads->ln_q_b), nullptr, stream));
  TP_TRY(launch_ln_bwd(wb(B.dkh), sb(S.y_k), stats_k, w->ln_k_w, wb(B.dyk), ln_part + 3ll * kLnBlocks * kC, R, G(grads->ln_k_w),
                       G(grads->ln_k_b), G(grads->k_proj_2_b), stream));
  TP_TRY(launch_ln_bwd(wb(B.dvh), sb(S.y_v), stats_v, w->ln_v_w, wb(B.dyv), ln_part + 6ll * kLnBlocks * kC, R, G(grads->ln_v_w),
                       G(grads->ln_v_b), G(grads->v_proj_2_b), stream));
  // ---- q_proj_1, k_proj_1.2, v_proj_1.2
  {
    GemmItem gi[5];
    int n = 0;
    const bool sk = wgrad_split(wb(B.dyk), kC, sb(S.h_kv), 2 * kC, R, 0, &gi[0]) && wgrad_split(wb(B.dyv), kC, sb(S.h_kv) + kC, 2 * kC, R, 1, &gi[1]);
    if (sk) {
      n = 2;
    } else {
      TP_TRY(wgrad(wb(B.dyk), kC, sb(S.h_kv), 2 * kC, R, kC, kC, G(grads->k_proj_2_w), kC, 1.0f, B.dyk_t, B.hkv_t, &gi[n++]));
      TP_TRY(wgrad(wb(B.dyv), kC, sb(S.h_kv) + kC, 2 * kC, R, kC, kC, G(grads->v_proj_2_w), kC, 1.0f, B.dyv_t, B.hkv_t + static_cast<size_t>(kC) * Rp * 2, &gi[n++]));
    }
    TP_TRY(dgrad(wb(B.dyk), kC, w->k_proj_2_w, kC, wb(B.dzkv), 2 * kC, R, kC, kC, 1.0f, B.w_k2t, &gi[n++]));        // dh_k  -> dzkv[:, :1024]
    TP_TRY(dgrad(wb(B.dyv), kC, w->v_proj_2_w, kC, wb(B.dzkv) + kC, 2 * kC, R, kC, kC, 1.0f, B.w_v2t, &gi[n++]));   // dh_v  -> dzkv[:, 1024:]
    TP_TRY(wgrad(wb(B.dyq), kC, sb(S.q), kC, Q, kC, kC, G(grads->q_proj_w), kC, 1.0f, B.dyq_t, B.q_t, &gi[n++]));
    TP_TRY(launch_gemms(gi, n, dev.sms, stream));
    if (sk) {
      TP_TRY(wgrad_reduce(0, 1.0f, G(grads->k_proj_2_w)));
      TP_TRY(wgrad_reduce(1, 1.0f, G(grads->v_proj_2_w)));
    }
  }
  // ---- k_proj_1.0 / v_proj_1.0:  z = xm W0^T + b   (no gradient to xm: the CLIP tower is frozen)
  TP_TRY(gelu_bwd_bias(wb(B.dzkv), sb(S.z_kv), 2 * kC, R, 2 * kC, G(grads->k_proj_0_b), G(grads->v_proj_0_b)));     // dz_kv + both bias gradients
  {
    GemmItem gi[2];
    TP_TRY(wgrad(wb(B.dzkv), 2 * kC, xm, kCm, R, kC, kCm, G(grads->k_proj_0_w), kCm, 1.0f, B.dzkv_t, B.xm_t, &gi[0]));
    TP_TRY(wgrad(wb(B.dzkv) + kC, 2 * kC, xm, kCm, R, kC, kCm, G(grads->v_proj_0_w), kCm, 1.0f, B.dzkv_t + static_cast<size_t>(kC) * Rp * 2, B.xm_t, &gi[1]));
    TP_TRY(launch_gemms(gi, 2, dev.sms, stream));
  }
  return TP_OK;
}
