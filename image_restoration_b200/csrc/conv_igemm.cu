// Host side of the implicit-GEMM convolution: descriptor validation, TMA tensor maps, tile / ring sizing, epilogue
// profile selection and the launch through the per-profile translation units (conv_epi*.cu).
// Device code: conv_common.cuh.  Reference semantics: include/b200ir.h (b200ir_conv_igemm).
#include "conv_common.cuh"

namespace b200ir {
extern template int launch_conv_variant<-1>(const ConvParams&, int, bool, int, int, int, cudaStream_t);
extern template int launch_conv_variant<0>(const ConvParams&, int, bool, int, int, int, cudaStream_t);
extern template int launch_conv_variant<1>(const ConvParams&, int, bool, int, int, int, cudaStream_t);
extern template int launch_conv_variant<2>(const ConvParams&, int, bool, int, int, int, cudaStream_t);
extern template int launch_conv_variant<3>(const ConvParams&, int, bool, int, int, int, cudaStream_t);
extern template int launch_conv_variant<4>(const ConvParams&, int, bool, int, int, int, cudaStream_t);
extern template int launch_conv_variant<5>(const ConvParams&, int, bool, int, int, int, cudaStream_t);
extern template int launch_conv_variant<6>(const ConvParams&, int, bool, int, int, int, cudaStream_t);

static int launch_conv(const ConvParams& p, bool row, int grid, int smem_bytes, int smem_max, cudaStream_t st) {
  switch (p.epi) {
    case 0: return launch_conv_variant<0>(p, p.block_k, row, grid, smem_bytes, smem_max, st);
    case 1: return launch_conv_variant<1>(p, p.block_k, row, grid, smem_bytes, smem_max, st);
    case 2: return launch_conv_variant<2>(p, p.block_k, row, grid, smem_bytes, smem_max, st);
    case 3: return launch_conv_variant<3>(p, p.block_k, row, grid, smem_bytes, smem_max, st);
    case 4: return launch_conv_variant<4>(p, p.block_k, row, grid, smem_bytes, smem_max, st);
    case 5: return launch_conv_variant<5>(p, p.block_k, row, grid, smem_bytes, smem_max, st);
    case 6: return launch_conv_variant<6>(p, p.block_k, row, grid, smem_bytes, smem_max, st);
    default: return launch_conv_variant<-1>(p, p.block_k, row, grid, smem_bytes, smem_max, st);
  }
}
}  // namespace b200ir

namespace b200ir {

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || sym == nullptr) {
    set_error("cuTensorMapEncodeTiled unavailable: %s", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

int encode_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
               const cuuint32_t* box, CUtensorMapSwizzle swz, const char* what) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return 1;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(ptr), dims, strides_b, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d (dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)", what,
              (int)r, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return 1;
  }
  return 0;
}

static int g_num_sms = 0;
static int g_smem_optin = 0;

static int init_device_info() {
  if (g_num_sms) return 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device");
    return 1;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    set_error("cudaGetDeviceProperties failed");
    return 1;
  }
  if (prop.major != 10) {
    set_error("device is sm_%d%d; libb200ir needs sm_100 (B200) and has no fallback", prop.major, prop.minor);
    return 1;
  }
  g_num_sms = prop.multiProcessorCount;
  g_smem_optin = (int)prop.sharedMemPerBlockOptin;
  return 0;
}

int device_check_impl() { return init_device_info(); }
int num_sms() { return init_device_info() ? 0 : g_num_sms; }
int smem_optin() { return init_device_info() ? 0 : g_smem_optin; }

}  // namespace b200ir

using namespace b200ir;

// Everything a launch needs besides the stream: kernel parameters (tensor maps included), kernel variant, grid and
// dynamic shared memory.  Built once per descriptor by build_conv_launch; b200ir_conv_plan keeps it.
struct ConvLaunch {
  ConvParams p;
  bool row;
  int grid, smem_bytes;
};

static int build_conv_launch(const b200ir_conv_desc* d, ConvLaunch& L) {
  B200IR_REQUIRE(d != nullptr, "conv_igemm: null desc");
  if (init_device_info()) return 1;
  B200IR_REQUIRE(d->tile_w > 0 && d->tile_h > 0 && d->tile_b > 0 && d->tile_w * d->tile_h * d->tile_b == kBlockM,
                 "conv_igemm: tile %dx%dx%d must cover exactly 128 positions", d->tile_b, d->tile_h, d->tile_w);
  B200IR_REQUIRE(d->tile_w <= 256 && d->tile_h <= 256 && d->tile_b <= 256, "conv_igemm: tile extent > 256");
  B200IR_REQUIRE(d->cin >= 16 && d->cin % 16 == 0, "conv_igemm: cin=%d must be a multiple of 16", d->cin);
  B200IR_REQUIRE(d->block_n >= 16 && d->block_n <= 256 && d->block_n % 16 == 0 && d->cout % d->block_n == 0,
                 "conv_igemm: block_n=%d invalid for cout=%d", d->block_n, d->cout);
  B200IR_REQUIRE(d->num_taps >= 1 && d->num_taps <= B200IR_MAX_TAPS, "conv_igemm: num_taps=%d", d->num_taps);
  B200IR_REQUIRE(d->num_views >= 1 && d->num_views <= B200IR_MAX_VIEWS, "conv_igemm: num_views=%d", d->num_views);
  B200IR_REQUIRE((d->out != nullptr || d->no_store) && d->weight != nullptr, "conv_igemm: null out/weight");
  B200IR_REQUIRE(d->m_w > 0 && d->m_h > 0 && d->m_b > 0, "conv_igemm: empty M extents");
  B200IR_REQUIRE((d->out_c_off % 8) == 0 && (d->out_stride_x % 8) == 0, "conv_igemm: output not 16B aligned");
  B200IR_REQUIRE(d->res_mode >= 0 && d->res_mode <= 2, "conv_igemm: res_mode");
  B200IR_REQUIRE(d->res_mode == 0 || d->res != nullptr, "conv_igemm: res_mode set but res is NULL");
  B200IR_REQUIRE(d->noise == nullptr || d->noise_gain != nullptr, "conv_igemm: noise without noise_gain");
  B200IR_REQUIRE(d->act >= 0 && d->act <= 2, "conv_igemm: act=%d", d->act);
  {
    const int ps_c = d->ps_c ? d->ps_c : d->block_n;
    B200IR_REQUIRE(d->ps_r == 0 || ((d->ps_r == 2 || d->ps_r == 3) && ps_c * d->ps_r * d->ps_r == d->cout &&
                                    ps_c % 16 == 0 && d->rgb_w == nullptr && d->noise == nullptr &&
                                    (d->res_mode == 0 || d->corr_top != nullptr)),
                   "conv_igemm: ps_r=%d needs ps_c (or block_n) = cout / ps_r^2 and a plain epilogue", d->ps_r);
    B200IR_REQUIRE(!d->use_tap_mask || d->cout / d->block_n <= 8, "conv_igemm: tap masks cover at most 8 N-tiles");
    B200IR_REQUIRE(d->demod_c == 0 || d->demod_c == d->cout || d->demod_c % d->block_n == 0 ||
                       d->tile_b * d->block_n <= kDemodTable,
                   "conv_igemm: demod_c=%d needs block_n | demod_c or a tile whose demod table fits shared memory",
                   d->demod_c);
  }
  B200IR_REQUIRE(d->rgb_w == nullptr || (d->rgb_part != nullptr && d->rgb_w_px > 0 && d->rgb_h > 0),
                 "conv_igemm: rgb_w needs rgb_part and the plane extents");
  B200IR_REQUIRE(!d->no_store || d->rgb_w != nullptr, "conv_igemm: no_store without a fused ToRGB leaves no output");

  ConvParams& p = L.p;
  memset(&p, 0, sizeof(p));
  p.block_k = (d->cin % 64 == 0) ? 64 : ((d->cin % 32 == 0) ? 32 : 16);
  p.k_chunks = d->cin / p.block_k;
  const CUtensorMapSwizzle swz = (p.block_k == 64)   ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (p.block_k == 32) ? CU_TENSOR_MAP_SWIZZLE_64B
                                                     : CU_TENSOR_MAP_SWIZZLE_32B;
  for (int v = 0; v < B200IR_MAX_VIEWS; ++v) {
    const b200ir_view& a = d->a[v < d->num_views ? v : 0];
    B200IR_REQUIRE(a.ptr != nullptr && a.c >= d->cin, "conv_igemm: view %d invalid", v);
    B200IR_REQUIRE((reinterpret_cast<uintptr_t>(a.ptr) & 15) == 0 && a.stride_w % 8 == 0 && a.stride_h % 8 == 0 &&
                       a.stride_b % 8 == 0,
                   "conv_igemm: view %d not 16B aligned", v);
    cuuint64_t dims[4] = {(cuuint64_t)a.c, (cuuint64_t)a.w, (cuuint64_t)a.h, (cuuint64_t)a.b};
    cuuint64_t strides[3] = {(cuuint64_t)a.stride_w * 2, (cuuint64_t)a.stride_h * 2, (cuuint64_t)a.stride_b * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.block_k, (cuuint32_t)d->tile_w, (cuuint32_t)d->tile_h, (cuuint32_t)d->tile_b};
    if (encode_map(&p.tmap_a[v], a.ptr, 4, dims, strides, box, swz, "activation")) return 1;
  }
  B200IR_REQUIRE(!d->w_per_image || (d->tile_b == 1 && d->row_mode == 0),
                 "conv_igemm: per-image weights need tile_b = 1 and the generic kernel");
  {
    const int k_total = d->num_taps * d->cin;
    cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)d->cout * (d->w_per_image ? (cuuint64_t)d->m_b : 1u)};
    cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.block_k, (cuuint32_t)d->block_n};
    if (encode_map(&p.tmap_b, d->weight, 2, dims, strides, box, swz, "weight")) return 1;
  }
  p.tile_w = d->tile_w; p.tile_h = d->tile_h; p.tile_b = d->tile_b;
  p.tiles_w = (d->m_w + d->tile_w - 1) / d->tile_w;
  p.tiles_h = (d->m_h + d->tile_h - 1) / d->tile_h;
  p.tiles_b = (d->m_b + d->tile_b - 1) / d->tile_b;
  p.tiles_n = d->cout / d->block_n;
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_b * p.tiles_n;
  p.block_n = d->block_n;
  p.num_taps = d->num_taps;
  p.m_w = d->m_w; p.m_h = d->m_h; p.m_b = d->m_b;
  for (int t = 0; t < d->num_taps; ++t) {
    B200IR_REQUIRE(d->tap_view[t] >= 0 && d->tap_view[t] < d->num_views, "conv_igemm: tap %d view", t);
    p.tap_view[t] = d->tap_view[t]; p.tap_dx[t] = d->tap_dx[t]; p.tap_dy[t] = d->tap_dy[t];
  }
  p.idesc = make_idesc_f16(kBlockM, d->block_n, false);
  p.smem_demod = (d->demod != nullptr && d->tile_b * d->block_n <= kDemodTable) ? 1 : 0;
  p.smem_aux = ((d->out_scale != nullptr || d->rgb_w != nullptr) && 5 * d->tile_b * d->block_n <= kDemodTable) ? 1 : 0;
  p.st256 = (!d->out_fp32 && d->out_c_off % 16 == 0 && d->out_stride_x % 16 == 0 && d->out_stride_y % 16 == 0 &&
             d->out_stride_b % 16 == 0 && (reinterpret_cast<uintptr_t>(d->out) & 31) == 0) ? 1 : 0;
  B200IR_REQUIRE(d->cout <= kMaxBias || d->bias == nullptr || !d->act || (d->corr_top != nullptr && d->ps_c <= kMaxBias),
                 "conv_igemm: cout=%d > %d needs a bias vector and no activation", d->cout, kMaxBias);
  // accumulator ring depth: the epilogue latency of a tile is hidden behind the main loops of the next
  // (acc_stages - 1) tiles; small tiles (short main loops) need a deeper ring
  int acc_stages = 2, acc_shift = 1;
  while (acc_stages < kMaxAccStages && 2 * acc_stages * d->block_n <= 512) {
    acc_stages *= 2;
    ++acc_shift;
  }
  p.acc_stages = acc_stages;
  p.acc_shift = acc_shift;
  uint32_t cols = 32;
  while (cols < (uint32_t)(acc_stages * d->block_n)) cols <<= 1;
  p.tmem_cols = cols;

  const int row_bytes = p.block_k * 2;
  const int stage_bytes = kBlockM * row_bytes + d->block_n * row_bytes;
  const int tail = kTailBytes;
  int stages = (g_smem_optin - 1024 - tail) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  B200IR_REQUIRE(stages >= 2, "conv_igemm: not enough shared memory for 2 stages");
  int smem_bytes = stages * stage_bytes + tail + 1024;
  {
    // weights resident in shared memory: single N-tile layers re-read the same weight tiles for every M-tile
    // (folded ConvUpLayer 64 -> 4x32: 147 KB of weights per 144 KB of activations, measured L2-bound); keep them
    // when at least three activation stages still fit
    const int a_bytes = kBlockM * row_bytes;
    const int w_bytes = d->num_taps * p.k_chunks * d->block_n * row_bytes;
    const int st_res = (g_smem_optin - 1024 - tail - w_bytes) / a_bytes;
    static int no_res = -1;
    if (no_res < 0) no_res = (getenv("B200IR_NO_RESIDENT") != nullptr) ? 1 : 0;
    if (!no_res && d->block_n == d->cout && !d->use_tap_mask && st_res >= 3 && d->num_taps * p.k_chunks >= 2 &&
        !d->w_per_image) {
      p.b_resident = 1;
      stages = st_res > kMaxStages ? kMaxStages : st_res;
      smem_bytes = w_bytes + stages * a_bytes + tail + 1024;
    }
  }
  p.stages = stages;

  p.out = d->out; p.out_fp32 = d->out_fp32;
  p.out_sx = d->out_stride_x; p.out_sy = d->out_stride_y; p.out_sb = d->out_stride_b;
  p.out_c_off = d->out_c_off;
  p.out_x_mul = d->out_x_mul; p.out_x_off = d->out_x_off; p.out_y_mul = d->out_y_mul; p.out_y_off = d->out_y_off;
  p.cout = d->cout;
  p.bias = d->bias; p.demod = d->demod; p.noise = d->noise; p.noise_gain = d->noise_gain;
  p.noise_sb = d->noise_stride_b; p.noise_sy = d->noise_stride_y;
  p.act = d->act; p.act_gain = d->act == 1 ? 1.4142135623730951f : 1.f; p.res_mode = d->res_mode; p.res = reinterpret_cast<const __half*>(d->res);
  p.res_sx = d->res_stride_x; p.res_sy = d->res_stride_y; p.res_sb = d->res_stride_b;
  p.res_w = d->res_w; p.res_h = d->res_h; p.res_scale = d->res_scale;
  p.res_mul = d->res_mul != 0.f ? d->res_mul : d->res_scale;
  p.ps_r = d->ps_r;
  p.corr_top = d->corr_top; p.corr_bot = d->corr_bot; p.corr_left = d->corr_left; p.corr_right = d->corr_right;
  p.ps_c = d->ps_c ? d->ps_c : d->block_n;
  p.ps_shift = -1;
  for (int sh = 0; sh < 16; ++sh)
    if ((1 << sh) == p.ps_c) p.ps_shift = sh;
  p.demod_c = d->demod_c ? d->demod_c : d->cout;
  p.bias_c = (d->corr_top != nullptr && d->ps_c > 0) ? d->ps_c : d->cout;
  for (int t = 0; t < 8; ++t) p.tap_mask[t] = d->use_tap_mask ? d->tap_mask[t] : 0xffffffffu;
  {
    static int dbg = -1;
    if (dbg < 0) dbg = (getenv("B200IR_DBG_SKIP_EPI") != nullptr) ? atoi(getenv("B200IR_DBG_SKIP_EPI")) : 0;
    p.dbg_skip_epi = dbg;
    // B200IR_EPI_PIPE: 0 / 1 force the pipelined TMEM loads of the fast epilogues off / on (A/B switch, tools/time_plan_ops.py)
    static int pipe = -2;
    if (pipe == -2) pipe = (getenv("B200IR_EPI_PIPE") != nullptr) ? atoi(getenv("B200IR_EPI_PIPE")) : -1;
    p.epi_pipe = pipe >= 0 ? pipe : 1;
    static int wait_ns = -1;
    if (wait_ns < 0) wait_ns = (getenv("B200IR_EPI_WAIT_NS") != nullptr) ? atoi(getenv("B200IR_EPI_WAIT_NS")) : 0;
    p.epi_wait_ns = wait_ns;
    static int peek = -1;
    if (peek < 0) peek = (getenv("B200IR_ROW_PEEK") != nullptr) ? atoi(getenv("B200IR_ROW_PEEK")) : 1;
    p.row_peek = peek;
  }
  // ---- specialised epilogue selection
  {
    const bool fp16_fast = !d->out_fp32 && (d->no_store || p.st256) &&
                           (d->cout <= kMaxBias || d->bias == nullptr || (d->corr_top != nullptr && d->ps_c <= kMaxBias)) &&
                           p.block_k >= 32;
    int flags = 0;
    bool ok = fp16_fast;
    if (d->demod != nullptr) { flags |= F_DEMOD; ok = ok && p.smem_demod; }
    if (d->noise != nullptr) flags |= F_NOISE;
    if (d->res_mode == 1) flags |= F_RES1;
    if (d->res_mode == 2) flags |= F_RES2;
    if (d->rgb_w != nullptr) { flags |= F_RGB; ok = ok && p.smem_aux; }
    else if (d->out_scale != nullptr) ok = false;
    if (d->no_store) flags |= F_NOSTORE;
    if (d->corr_top != nullptr) {
      flags |= F_UPFOLD;
      ok = ok && d->ps_r == 2 && p.ps_shift >= 0 && d->res_mode == 2 && d->corr_bot && d->corr_left && d->corr_right;
    }
    p.epi = -1;
    for (int i = 0; ok && i < kNumEpiProfiles; ++i)
      if (epi_profile_flags(i) == flags) p.epi = i;
    static int force_generic = -1;
    if (force_generic < 0) force_generic = (getenv("B200IR_GENERIC_EPI") != nullptr) ? 1 : 0;
    if (force_generic) p.epi = -1;
    B200IR_REQUIRE(d->corr_top == nullptr || p.epi == 6,
                   "conv_igemm: the folded ConvUpLayer epilogue needs ps_r = 2, a power-of-two ps_c, res_mode = 2, fp16 "
                   "32-byte aligned output and all four correction buffers");
    p.slope = d->act == 1 ? 0.2f : (d->act == 2 ? d->act_slope : 1.f);
    if (p.epi >= 0 && d->res_mode != 0) p.act_gain *= d->res_scale;
    // 256-column tiles have two accumulator stages, i.e. two busy epilogue groups: the residual epilogues (global gathers
    // per chunk: latency-bound) drain each tile as two halves over all three groups.  Measured (B200IR_EPI_SPLIT=0/1 in one
    // box): folded ConvUpLayer 237 -> 219 us, 256 -> 256 + bilinear residual 192 -> 182, 64 -> 256 stride 2 100 -> 92;
    // the demodulation-only epilogues of the transposed convs get slower when split (250 -> 273 us), so they are not
    static int split_env = -1;
    if (split_env < 0) split_env = (getenv("B200IR_EPI_SPLIT") != nullptr) ? atoi(getenv("B200IR_EPI_SPLIT")) : 1;
    p.epi_split = (split_env && p.epi >= 0 && p.acc_stages == 2 && d->block_n == 256 && (flags & (F_RES1 | F_RES2)) &&
                   !(flags & F_RGB)) ? 1 : 0;
  }
  p.w_img_rows = d->w_per_image ? d->cout : 0;
  p.out_scale = d->out_scale; p.rgb_w = d->rgb_w; p.rgb_part = d->rgb_part; p.no_store = d->no_store;
  p.rgb_w_px = d->rgb_w_px; p.rgb_plane = (long long)d->rgb_h * d->rgb_w_px;
  p.rgb_image = (long long)d->m_b * 3 * p.rgb_plane;

  // ---- row mode eligibility: plain 3x3 stride-1 conv on one view, 128-pixel row tiles, weights fit in smem
  {
    bool row_ok = d->num_views == 1 && d->num_taps == 9 && d->tile_w == 128 && d->tile_h == 1 && d->tile_b == 1 &&
                  d->block_n == d->cout && d->out_x_mul == 1 && d->out_y_mul == 1 && d->out_x_off == 0 &&
                  d->out_y_off == 0 && d->row_mode != 0 &&
                  (d->cout == 16 || d->cout == 32 || d->cout == 64);  // N = 3*cout <= 256, ring of 512/cout slots
    for (int t = 0; row_ok && t < 9; ++t)
      row_ok = d->tap_view[t] == 0 && d->tap_dx[t] == (t % 3) - 1 && d->tap_dy[t] == (t / 3) - 1;
    if (row_ok) {
      const int w_bytes = 9 * p.k_chunks * d->block_n * row_bytes;
      const int slot_bytes = 136 * row_bytes;
      const int tail_r = kTailBytes;
      int slots = (g_smem_optin - 1024 - tail_r - w_bytes) / slot_bytes;
      if (slots > kMaxStages) slots = kMaxStages;
      // Two CTAs per SM for the leanest layers (cout <= 32 with the plain epilogue profiles, 60-72 registers): there
      // the single MMA-issuing thread is the bound (measured 32 -> 32: main-loop floor 98 us against 61 us of HBM
      // time, tensor pipe 20 % busy), and a second resident CTA is a second issuing thread.  Each CTA takes half the
      // shared memory and 256 TMEM columns (ring of 8 accumulators).
      static int dual_env = -1;
      if (dual_env < 0) dual_env = (getenv("B200IR_ROW_DUAL") != nullptr) ? atoi(getenv("B200IR_ROW_DUAL")) : 1;
      bool dual = dual_env && (p.epi == 0 || p.epi == 1) && d->cout <= 32 && p.k_chunks == 1 && d->max_ctas <= 0;
      if (dual) {
        const int budget = 233472 / 2 - 1024 - 512;  // per-SM shared memory / 2, minus the per-CTA reservation
        const int s2 = (budget - 1024 - tail_r - w_bytes) / slot_bytes;
        if (s2 >= 4) {
          slots = s2 > kMaxStages ? kMaxStages : s2;
        } else {
          dual = false;
        }
      }
      const int row_ctas = dual ? 2 * g_num_sms : g_num_sms;
      if (slots >= 2 * p.k_chunks) {
        // re-encode the activation map with the 130-pixel halo box
        const b200ir_view& a = d->a[0];
        cuuint64_t dims[4] = {(cuuint64_t)a.c, (cuuint64_t)a.w, (cuuint64_t)a.h, (cuuint64_t)a.b};
        cuuint64_t strides[3] = {(cuuint64_t)a.stride_w * 2, (cuuint64_t)a.stride_h * 2, (cuuint64_t)a.stride_b * 2};
        cuuint32_t box[4] = {(cuuint32_t)p.block_k, 130u, 1u, 1u};
        if (encode_map(&p.tmap_a[0], a.ptr, 4, dims, strides, box, swz, "activation(row)")) return 1;
        // Rows per work item: an item of R output rows costs R + 2 input rows (halo), and the CTA with the most items
        // sets the time, so pick the chunk count that minimises ceil(items / CTAs) * (R + 2) (measured at batch 64,
        // 128x384: 3 chunks of 43 rows -> 4 items per CTA = 180 row times, against 204 for 4 chunks of 32).
        int R = d->m_h;
        {
          const int ctas = (d->max_ctas > 0 && d->max_ctas < row_ctas) ? d->max_ctas : row_ctas;
          long long best = -1;
          for (int c = 1; c <= d->m_h; ++c) {
            const int r = (d->m_h + c - 1) / c;
            if (r < 4 && c > 1) break;
            const long long items = (long long)d->m_b * p.tiles_w * ((d->m_h + r - 1) / r);
            const long long cost = ((items + ctas - 1) / ctas) * (r + 2);
            if (best < 0 || cost < best) {
              best = cost;
              R = r;
            }
          }
        }
        p.row_R = R;
        p.row_chunks = (d->m_h + R - 1) / R;
        p.row_items = d->m_b * p.tiles_w * p.row_chunks;
        {
          // balanced alternative: equal row ranges per CTA, cut at (image, segment) boundaries; a CTA then pays its rows
          // plus two halo rows per (partial) pair it touches.  Taken when that is at least 3 % shorter than the uniform grid
          // (64 x 192 levels at batch 64: 128 whole-image items on 148 SMs -> 59 instead of 66 row times).
          static int bal_env = -1;
          if (bal_env < 0) bal_env = (getenv("B200IR_ROW_BALANCED") != nullptr) ? atoi(getenv("B200IR_ROW_BALANCED")) : 1;
          const int ctas = (d->max_ctas > 0 && d->max_ctas < row_ctas) ? d->max_ctas : row_ctas;
          const long long total = (long long)d->m_b * p.tiles_w * d->m_h;
          const long long rows_cta = (total + ctas - 1) / ctas;
          const long long pairs_cta = (rows_cta + d->m_h - 1) / d->m_h + 1;
          const long long cost_bal = rows_cta + 2 * pairs_cta;
          const long long items_u = (long long)p.row_items;
          const long long cost_uni = ((items_u + ctas - 1) / ctas) * (R + 2);
          p.row_balanced = (bal_env && total >= 4LL * ctas && cost_bal * 100 < cost_uni * 97) ? 1 : 0;
        }
        // too little parallelism at small batch: generic tiles are faster (row_mode == 2 forces the variant: tests)
        row_ok = (long long)d->m_b * p.tiles_w * d->m_h >= 16LL * g_num_sms || d->row_mode == 2;
        p.row_slots = slots;
        p.row_slot_bytes = slot_bytes;
        p.row_w_bytes = w_bytes;
        p.desc_mode = 0;
        for (int n = 1; n <= 3; ++n) p.idesc_n[n - 1] = make_idesc_f16(kBlockM, n * d->block_n, false);
        const int smem_row = w_bytes + slots * slot_bytes + tail_r + 1024;
        if (row_ok) {
          if (dual) {
            p.acc_stages = 256 / d->block_n;
            p.acc_shift = 0;
            while ((1 << p.acc_shift) < p.acc_stages) ++p.acc_shift;
            p.tmem_cols = 256;
          }
          int grid_r = (p.row_items < row_ctas && !p.row_balanced) ? p.row_items : row_ctas;
          if (d->max_ctas > 0 && grid_r > d->max_ctas) grid_r = d->max_ctas;
          L.row = true;
          L.grid = grid_r;
          L.smem_bytes = smem_row;
          return 0;
        }
        // not taken: restore the generic activation map (box = tile)
        cuuint32_t box_g[4] = {(cuuint32_t)p.block_k, (cuuint32_t)d->tile_w, (cuuint32_t)d->tile_h, (cuuint32_t)d->tile_b};
        if (encode_map(&p.tmap_a[0], a.ptr, 4, dims, strides, box_g, swz, "activation")) return 1;
      }
    }
  }

  int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
  if (d->max_ctas > 0 && grid > d->max_ctas) grid = d->max_ctas;
  // ---- CTA pairs (cta_group::2): streamed-weight layers with N-tiles of 128 / 256 columns and at least one work item (two
  // M-tiles of one N-tile) per pair of SMs -- with fewer, single CTAs spread the tiles over more SMs (measured: 256 -> 256 at
  // 8x24 / 4x12 lose 2 us as pairs, 512 -> 4x512 at 9x25 gains 12 us); B200IR_CTA_PAIR=0 keeps every layer on single CTAs
  {
    static int pair_env = -1;
    if (pair_env < 0) pair_env = (getenv("B200IR_CTA_PAIR") != nullptr) ? atoi(getenv("B200IR_CTA_PAIR")) : 1;
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
    if (pair_env && !p.b_resident && p.block_k == 64 && d->block_n >= 128 && d->block_n % 16 == 0 && !d->w_per_image &&
        d->max_ctas <= 0 && m_tiles >= 2 && g_num_sms >= 2 && ((m_tiles + 1) / 2) * p.tiles_n >= g_num_sms / 2) {
      const int k_total = d->num_taps * d->cin;
      cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)d->cout};
      cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
      cuuint32_t box[2] = {(cuuint32_t)p.block_k, (cuuint32_t)(d->block_n / 2)};
      if (encode_map(&p.tmap_b2, d->weight, 2, dims, strides, box, swz, "weight(pair)")) return 1;
      p.pair = 1;
      p.pair_tiles = ((m_tiles + 1) / 2) * p.tiles_n;
      p.idesc_pair = make_idesc_f16(2 * kBlockM, d->block_n, false);
      const int row_b = p.block_k * 2;
      const int stage_pair = kBlockM * row_b + (d->block_n / 2) * row_b;
      int st_pair = (g_smem_optin - 1024 - kTailBytes) / stage_pair;
      if (st_pair > kMaxStages) st_pair = kMaxStages;
      p.stages = st_pair;
      smem_bytes = st_pair * stage_pair + kTailBytes + 1024;
      const int pairs = g_num_sms / 2 < p.pair_tiles ? g_num_sms / 2 : p.pair_tiles;
      grid = 2 * pairs;
    }
  }
  L.row = false;
  L.grid = grid;
  L.smem_bytes = smem_bytes;
  return 0;
}

extern "C" int b200ir_conv_igemm(const b200ir_conv_desc* d, void* stream) {
  ConvLaunch L;
  if (build_conv_launch(d, L)) return 1;
  return launch_conv(L.p, L.row, L.grid, L.smem_bytes, g_smem_optin, reinterpret_cast<cudaStream_t>(stream));
}

// ---- launch plans: descriptor validation, tensor-map encoding and tile / ring sizing done once, launched many times
struct b200ir_conv_plan {
  ConvLaunch L;
  int device;
};

extern "C" int b200ir_conv_plan_create(const b200ir_conv_desc* d, b200ir_conv_plan** plan) {
  B200IR_REQUIRE(plan != nullptr, "conv_plan_create: null plan pointer");
  *plan = nullptr;
  b200ir_conv_plan* pl = static_cast<b200ir_conv_plan*>(aligned_alloc(64, (sizeof(b200ir_conv_plan) + 63) / 64 * 64));
  B200IR_REQUIRE(pl != nullptr, "conv_plan_create: out of host memory");
  if (build_conv_launch(d, pl->L)) {
    free(pl);
    return 1;
  }
  cudaGetDevice(&pl->device);
  *plan = pl;
  return 0;
}

extern "C" int b200ir_conv_plan_launch(const b200ir_conv_plan* plan, void* stream) {
  B200IR_REQUIRE(plan != nullptr, "conv_plan_launch: null plan");
  const ConvLaunch& L = plan->L;
  return launch_conv(L.p, L.row, L.grid, L.smem_bytes, g_smem_optin, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" void b200ir_conv_plan_destroy(b200ir_conv_plan* plan) { free(plan); }
