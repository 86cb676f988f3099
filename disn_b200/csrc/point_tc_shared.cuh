// Pieces of the tensor-core point kernel (point_tc.cu) that are also host-visible: tile geometry, accumulator
// column maps, operand-split stores.  Included inside namespace disn { namespace { ... } }.
#pragma once

constexpr int NX = 3;                 // activation (A operand) ring slots
constexpr int NG = 2;                 // gather ring slots
constexpr int W_TILE = 16384;         // 128 rows x 64 k x bf16 (one B tile half)
constexpr int W_STAGE = 2 * W_TILE;   // [W_hi | W_lo] for one (K-slice, N-block)
constexpr int X_HALF = 8192;          // 64 rows x 64 k x bf16
constexpr int G_LD = 65;              // padded point stride of the gather ring
constexpr int PTS = 64;               // points per CTA per tile
constexpr int NTHREADS = 512;
constexpr int FIRST_L0_POS = 57;        // position of the even stream's L0 stage in the per-tile consumption cycle
constexpr int STAGES_PER_STREAM = 33; // 1 + 8 + 16 + 8 weight stages (pair-level, 64 KB each: 2 CTA halves x [hi|lo])
// shared-memory table of small fp32 parameters per stream
constexpr int SB_B2 = 0, SB_B3 = 256, SB_B4 = 768, SB_B5 = 1280, SB_W6 = 1536, SB_W1 = 1792, SB_B1 = 1984, SB_STRIDE = 2048;
constexpr int XSLOTS_PER_STREAM = 20; // 4 + 8 + 8 activation slices drained from TMEM (layer-1 output has its own slot)


// feature index held by (half h, thread-column c) of an accumulator (2x2 datapath, N=256 per MMA)
__host__ __device__ constexpr int fout(int h, int c) { return (c / 128) * 256 + h * 128 + (c % 128); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// TMEM column base of the accumulator written by tensor-core layer `layer` (0..3) of a stream with parity q.
// Even streams: acc2 [128,256) acc3 [256,512) acc4 [0,256) acc5 [256,384); odd streams use the mirror image (^256), so the
// NEXT stream's first accumulator always has 128 free columns while the current stream's last two layers are resident.
__host__ __device__ constexpr uint32_t acc_col(int layer, int q) {
  return (layer == 0 ? 128u : (layer == 2 ? 0u : 256u)) ^ ((uint32_t)q << 8);
}

struct TileCoord { int b; int64_t n0; };
__device__ __forceinline__ TileCoord tile_coord(int64_t tile, int64_t tiles_per_img) {
  TileCoord t;
  t.b = (int)(tile / tiles_per_img);
  t.n0 = (tile % tiles_per_img) * (2 * PTS);
  return t;
}

// operand-format modes of the kernel
constexpr int MODE_BF16X3 = 0;   // x = hi + lo (bf16): hi*hi + lo*hi + hi*lo, 12 kind::f16 MMAs per 64-wide K slice
constexpr int MODE_F16F8 = 1;    // fp16 main product + two e5m2 correction products (kind::f8f6f4, twice the rate):
                                 //   a.w ~= h(a).h(w) + e((a-h(a)).2^s1).e(w.2^-s1) + e(a.2^-s2).e((w-h(w)).2^s2), 4 + 2 + 2 MMAs
constexpr int X8_TILE = 4096;    // 64 rows x 64 k x 1 B (SW64)
constexpr int W8_TILE = 8192;    // 128 rows x 64 k x 1 B (SW64)

// write one thread's 32 consecutive K values of row p into an A-tile slot.
// MODE_BF16X3: x0 = bf16 hi tile, x1 = bf16 lo tile (both SW128).
// MODE_F16F8 : x0 = fp16 tile (SW128), x1 = [e5m2((a-h).sc_lo) | e5m2(a.sc_hi)] two SW64 byte tiles.
// `res` / `copy`: which of the two e5m2 tiles the consuming layer uses (a layer may drop one or both corrections).
// MODE_F16F8 also folds the slice's fp16 values into `amax` (all values are post-ReLU, i.e. >= 0): an activation above
// fp16's 65504 becomes +inf there, which the kernel reports through PointJob::status instead of producing a silent inf.
template <int kMode>
__device__ __forceinline__ void store_slice(uint8_t* x0, uint8_t* x1, int p, int h, const float* v, float sc_lo,
                                            float sc_hi, __half2& amax, bool res = true, bool copy = true) {
  if constexpr (kMode == MODE_BF16X3) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) tc::split_bf16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t off = tc::sw128_offset((uint32_t)p, (uint32_t)(h * 4 + q));
      *reinterpret_cast<uint4*>(x0 + off) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
      *reinterpret_cast<uint4*>(x1 + off) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
    }
  } else {
    uint32_t m[16], lo[8], hi[8];
    const __half2 sc_hi2 = __float2half2_rn(sc_hi);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = v[2 * j], b = v[2 * j + 1];
      const __half2 hh = __floats2half2_rn(a, b);
      amax = __hmax2(amax, hh);
      m[j] = *reinterpret_cast<const uint32_t*>(&hh);
      if (res) {     // e5m2 of the fp16 rounding residual (first correction's A operand)
        const float ra = a - __low2float(hh), rb = b - __high2float(hh);
        const uint32_t l = __nv_cvt_float2_to_fp8x2(make_float2(ra * sc_lo, rb * sc_lo), __NV_SATFINITE, __NV_E5M2);
        if (j & 1) lo[j >> 1] |= l << 16;
        else lo[j >> 1] = l;
      }
    }
    if (copy) {      // e5m2 copy of a (second correction's A operand); skipped for a layer that drops that correction
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const __half2 hh = *reinterpret_cast<const __half2*>(&m[j]);
        const __half2 hs = __hmul2(hh, sc_hi2);    // power-of-two scale: exact up to fp16 underflow (below e5m2 precision)
        const uint32_t g = __nv_cvt_halfraw2_to_fp8x2(*reinterpret_cast<const __half2_raw*>(&hs), __NV_SATFINITE, __NV_E5M2);
        if (j & 1) hi[j >> 1] |= g << 16;
        else hi[j >> 1] = g;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<uint4*>(x0 + tc::sw128_offset((uint32_t)p, (uint32_t)(h * 4 + q))) =
          make_uint4(m[4 * q], m[4 * q + 1], m[4 * q + 2], m[4 * q + 3]);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const uint32_t off = tc::sw64_offset((uint32_t)p, (uint32_t)(h * 2 + q));
      if (res) *reinterpret_cast<uint4*>(x1 + off) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
      if (copy) *reinterpret_cast<uint4*>(x1 + X8_TILE + off) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
    }
  }
}

