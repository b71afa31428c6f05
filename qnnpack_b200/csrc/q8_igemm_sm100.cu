// q8gemm / q8conv for sm_100a: persistent, warp-specialised implicit GEMM on the 5th-gen tensor cores.
//
// Replaces (reference, paths relative to its root):
//   src/operator-run.c:770-804 + :39-70   gemm case  -> q8gemm_ukernel_4x4c2__sse2 (src/q8gemm/4x4c2-sse2.c:14-318)
//   src/operator-run.c:805-844 + :183-217 conv case  -> q8conv_ukernel_4x4c2__sse2 (src/q8conv/4x4c2-sse2.c:14-273)
//   src/indirection.c:18-79 (the pointer table is never built: tap -> address is computed in the load stage)
//   the Q31 epilogue those micro-kernels inline (src/q8gemm/4x4c2-sse2.c:178-278)
//
// Arithmetic (bit-exact by construction, all int32, order independent).  Two regroupings of the
// reference's  acc = bias' + sum_k a_k (w_k - kzp),  bias' = b + K*izp*kzp - izp*sum_k w,  padded taps reading izp:
//   "folded": every additive term runs on the tensor core — u8 x s8 UMMAs with B1 = w XOR 0x80, a constant
//             (128 - kzp) operand, and bias' as base-255 digits against a constant A row; the epilogue only requantises;
//   "ones"  : the reference's own XZP algebra (src/q8gemm/4x8c2-xzp-neon.c:26-67, pack.h:216-232): raw u8 x u8 UMMA,
//             sum_k a[m][k] from one extra B row of ones, epilogue adds bias'[n] - kzp * rowsum[m].
//
// Structure (one CTA per SM, 768 threads):
//   warps 0-15   epilogue, two pairs of 8 warps alternating work items: tcgen05.ld -> Q31 requantisation ("U" form,
//                requant_math.h) -> saturating pack -> the pair's smem image of the output tile (or direct stores)
//   warps 16-18  TMEM allocation (warp 16) + UMMA issue, one lane each, sub-tiles split round-robin
//   warps 19-22  loaders: TMA (1x1 / FC), cp.async (general conv: tap -> address in the load stage), or the raw-row
//                ring + branch-free K-row builder (3x3 over 3 channels); canonical K-major no-swizzle layout
//                [sub-tile][k-chunk][row][16 B]
//   warp  23     lanes 0/1: one bulk (1-D TMA) store per finished item of pair 0/1
// A work item is `mt` (<= 8) consecutive 128-row sub-tiles x one n-tile: the sub-tiles' accumulators sit side
// by side in one TMEM stage (mt * n_mma <= 256 columns), which amortises every per-item synchronisation
// over up to 1024 rows — essential for the narrow (N = 16..96) projection layers.
// Pipelines: smem ring full/empty mbarriers (loaders <-> UMMA), 2-4 TMEM accumulator stages full/empty
// (UMMA <-> epilogue pairs), out_full/out_free (epilogue pair <-> store lane).  int32 accumulators never leave
// TMEM/registers.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#include <type_traits>

#include "q8_igemm_sm100.cuh"
#include "requant_dev.cuh"
#include "sm100_ptx.cuh"

namespace q8 {

constexpr int kEpiWarps = 16;
constexpr int kMmaWarp = kEpiWarps;   // first UMMA-issuing warp (it also owns the TMEM allocation)
// UMMA-issuing warps, one lane each; warp w issues for the sub-tiles j = w, w + kMmaWarps, ... of every item.  A single
// issuing thread was the bottleneck of the write-heavy layers: nvcc wraps each tcgen05.mma in a 15-30 instruction
// elect/broadcast sequence and the thread shares its scheduler with four epilogue warps (see also the depthwise kernel).
constexpr int kMmaWarps = 3;  // 16 + 3 + 4 + 1 = 24 warps = 768 threads keeps the 80-register budget of the epilogue
constexpr int kLoadWarp0 = kMmaWarp + kMmaWarps;
constexpr int kLoadWarps = 4;
constexpr int kLoadThreads = kLoadWarps * 32;
constexpr int kStoreWarp = kLoadWarp0 + kLoadWarps;       // lanes 0/1: bulk-store issue for epilogue pair 0/1
constexpr int kThreads = (kStoreWarp + 1) * 32;           // 768
constexpr int kEpiPairThreads = 256;
constexpr int kTmemCols = 512;
constexpr int kMaxRawBufs = 4;

struct __align__(8) SmemCtl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[kMaxAccStages];
  uint64_t tmem_empty[kMaxAccStages];
  uint64_t b_full;
  uint64_t out_full[4];  // [pair * 2 + buffer] epilogue pair -> store thread: the staged output tile is complete (256 arrivals)
  uint64_t out_free[4];  // [pair * 2 + buffer] store thread -> epilogue pair: the staging buffer may be overwritten
  uint64_t raw_full[kMaxRawBufs];  // raw-row staging (3x3x3 stem loader): bulk copies landed / 128 loader threads done
  uint64_t raw_empty[kMaxRawBufs];
  uint32_t tmem_base;
};

struct Item {
  long long m0;  // first row of the item
  int g, nt;
  int mt_eff;    // sub-tiles that contain at least one valid row
};

// Work item -> (group, super-tile, n-tile).  Every role decodes every item, the single UMMA-issuing thread among
// them, so this must be cheap: 32-bit arithmetic only (the host guarantees total_items, m_tiles < 2^31) and no
// division at all in the common single-n-tile, single-group case.  (64-bit div/mod here cost ~1 us per item.)
__device__ __forceinline__ Item decode_item(const IgemmParams& p, long long item64) {
  const uint32_t item = (uint32_t) item64;
  Item it;
  uint32_t st;
  if (p.n_tiles == 1 && p.groups == 1) {
    it.nt = 0;
    it.g = 0;
    st = item;
  } else {
    const uint32_t n_tiles = (uint32_t) p.n_tiles, m_super = (uint32_t) p.m_super;
    const uint32_t rest = item / n_tiles;
    it.nt = (int) (item - rest * n_tiles);
    it.g = (int) (rest / m_super);
    st = rest - (uint32_t) it.g * m_super;
  }
  it.m0 = (long long) st * (p.mt * kTileM);
  const uint32_t left = (uint32_t) p.m_tiles - st * (uint32_t) p.mt;
  it.mt_eff = left < (uint32_t) p.mt ? (int) left : p.mt;
  return it;
}

// ------------------------------------------------------------------------------------------------
// loaders
// ------------------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void copy_piece(uint32_t dst, const uint8_t* src) {
  if constexpr (VEC >= 4) {
    cp_async<VEC>(dst, src);
  } else {
    const uint8_t v = __ldg(src);
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(dst), "r"((uint32_t) v) : "memory");
  }
}

template <int VEC>
__device__ __forceinline__ void fill_piece(uint32_t dst, uint32_t byte4) {
  if constexpr (VEC == 16) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(byte4) : "memory");
  } else if constexpr (VEC == 8) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %1};" ::"r"(dst), "r"(byte4) : "memory");
  } else if constexpr (VEC == 4) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"(byte4) : "memory");
  } else {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(dst), "r"(byte4 & 0xFFu) : "memory");
  }
}

// 1x1 / fully-connected: row m of A is `gic` contiguous bytes at in + m*in_stride + g*gic.
// Lane mapping: 8 rows x 4 pieces per warp pass, so that the 8 lanes of each shared-memory store
// phase hit 8 different rows (distinct banks) while each row still reads 4*VEC contiguous bytes.
template <int VEC>
__device__ __forceinline__ void load_a_gemm(const IgemmParams& p, const Item& it, int ks, uint32_t a_stage, int ltid) {
  const int k0 = ks * p.skc * 16;
  int k1 = k0 + p.skc * 16;
  k1 = k1 < p.K ? k1 : p.K;
  const int pps = (k1 - k0) / VEC;  // pieces per row in this stage
  const int lane = ltid & 31, lw = ltid >> 5;
  const int rsub = lane & 7, psub = lane >> 3;
  const uint8_t* base = p.in + (size_t) it.g * p.gic + k0;
  const int rows = it.mt_eff * kTileM;
  const uint32_t sub_bytes = (uint32_t) p.skc * kChunkBytes;
#pragma unroll 1
  for (int rr = lw * 8 + rsub; rr < rows; rr += kLoadWarps * 8) {
    const long long m = it.m0 + rr;
    if (m < p.M) {
      const uint8_t* src = base + (size_t) m * p.in_stride;
      const uint32_t drow = a_stage + (uint32_t) (rr >> 7) * sub_bytes + (uint32_t) (rr & 127) * 16;
#pragma unroll 2
      for (int pc = psub; pc < pps; pc += 4) {
        const int kr = pc * VEC;
        copy_piece<VEC>(drow + (kr >> 4) * kChunkBytes + (kr & 15), src + kr);
      }
    }
  }
}

// generic convolution: thread = output pixel (row).  K index k = tap*gic + c, tap = ky*kw + kx; the
// tap's input pixel is ((n*H + iy)*W + ix) with the reference's unsigned bounds test
// (src/indirection.c:56-63); out-of-bounds taps are filled with the byte izp (src/convolution.c:336).
template <int VEC>
__device__ __forceinline__ void load_a_conv(const IgemmParams& p, const Item& it, int ks, uint32_t a_stage, int ltid) {
  const int k0 = ks * p.skc * 16;
  int k1 = k0 + p.skc * 16;
  k1 = k1 < p.K ? k1 : p.K;
  const int tap0 = k0 / p.gic, c0 = k0 % p.gic;
  const int ky0 = tap0 / p.kw, kx0 = tap0 % p.kw;
  const uint32_t fill = (uint32_t) p.izp * 0x01010101u;
  for (int j = 0; j < it.mt_eff; j++) {
    const long long m = it.m0 + (long long) j * kTileM + ltid;
    if (m < p.M) {  // (no early exit: every loader thread must still arrive on the stage barrier)
      const uint32_t mu = (uint32_t) m;  // M < 2^31 (host-checked): 32-bit divisions only
      const uint32_t t = mu / (uint32_t) p.out_w;
      const int ox = (int) (mu - t * (uint32_t) p.out_w);
      const uint32_t n = t / (uint32_t) p.out_h;
      const int oy = (int) (t - n * (uint32_t) p.out_h);
      const int iy0 = oy * p.stride_h - p.pad_top, ix0 = ox * p.stride_w - p.pad_left;
      int c = c0, ky = ky0, kx = kx0;
      const uint32_t drow = a_stage + (uint32_t) (j * p.skc) * kChunkBytes + (uint32_t) ltid * 16;
      const uint8_t* img = p.in + (size_t) n * p.in_h * p.in_w * p.in_stride + (size_t) it.g * p.gic;
      for (int kr = 0; kr < k1 - k0; kr += VEC) {
        const int iy = iy0 + ky * p.dil_h, ix = ix0 + kx * p.dil_w;
        const uint32_t dst = drow + (kr >> 4) * kChunkBytes + (kr & 15);
        if ((unsigned) iy < (unsigned) p.in_h && (unsigned) ix < (unsigned) p.in_w) {
          copy_piece<VEC>(dst, img + ((size_t) iy * p.in_w + ix) * p.in_stride + c);
        } else {
          fill_piece<VEC>(dst, fill);
        }
        c += VEC;
        if (c >= p.gic) {
          c = 0;
          if (++kx == p.kw) {
            kx = 0;
            ++ky;
          }
        }
      }
    }
  }
}

// Small-channel 3x3 convolution (the MobileNetV2 stem: 3 input channels, K = 27): with dense pixels and
// dilation_w == 1 the three taps of a kernel row are ONE contiguous run of kw*gic = 9 bytes, at an arbitrary byte
// alignment.  A thread gathers its pixel's three runs with 3 aligned 32-bit loads each (funnel-shifted into
// place), packs the 27 bytes of the K row in registers and writes them with two 16-byte shared stores.  Up to
// four pixels (sub-tiles) are in flight per thread so that the loads overlap.  Taps outside the image read izp.
constexpr int kRun = 9;

__device__ __forceinline__ void run9_load(const uint8_t* src, uint32_t (&w)[3]) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(src) & 3;
  const uint32_t* base = reinterpret_cast<const uint32_t*>(src - a);
  w[0] = __ldg(base);
  w[1] = __ldg(base + 1);
  w[2] = __ldg(base + 2);
}

// Raw-row staging for the run loader: the input rows an item needs are one or two CONTIGUOUS byte ranges of the NHWC
// tensor (an item may end in one image and continue in the next), so thread 0 fetches them with bulk copies and the 128
// loader threads then build the K rows from shared memory instead of waiting on ~200 scattered global loads each.
struct RawSeg {
  const uint8_t* g;   // 16-byte aligned start of the copy
  uint32_t bytes;     // multiple of 16
  uint32_t soff;      // offset of the copy inside the raw buffer
  int n, iy_lo;       // image and first input row it covers
  uint32_t delta;     // bytes between the aligned start and the first byte of row iy_lo
};

__device__ __forceinline__ int raw_segments(const IgemmParams& p, const Item& it, RawSeg (&seg)[2]) {
  const long long m_end = min(it.m0 + (long long) it.mt_eff * kTileM, p.M);
  const uint32_t ma = (uint32_t) it.m0, mb = (uint32_t) (m_end - 1);
  const uint32_t ta = ma / (uint32_t) p.out_w, tb = mb / (uint32_t) p.out_w;
  const int na = (int) (ta / (uint32_t) p.out_h), nb = (int) (tb / (uint32_t) p.out_h);
  const int oya = (int) (ta - (uint32_t) na * p.out_h), oyb = (int) (tb - (uint32_t) nb * p.out_h);
  const size_t row_bytes = (size_t) p.in_w * 3;
  const uint8_t* tensor_end = p.in + (size_t) p.raw_batch * p.in_h * row_bytes;
  const int nseg = na == nb ? 1 : 2;
  uint32_t soff = 0;
  for (int s = 0; s < nseg; s++) {
    const int n = s == 0 ? na : nb;
    int lo = (s == 0 ? oya * p.stride_h - p.pad_top : 0);
    int hi = (s == nseg - 1 ? oyb * p.stride_h - p.pad_top + (p.kh - 1) * p.dil_h : p.in_h - 1);
    lo = lo < 0 ? 0 : lo;
    hi = hi > p.in_h - 1 ? p.in_h - 1 : hi;
    const uint8_t* g0 = p.in + ((size_t) n * p.in_h + lo) * row_bytes;
    const uint8_t* g1 = p.in + ((size_t) n * p.in_h + hi + 1) * row_bytes;
    const uint8_t* ga = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(g0) & ~(uintptr_t) 15);
    const uint8_t* gb = reinterpret_cast<const uint8_t*>((reinterpret_cast<uintptr_t>(g1) + 15) & ~(uintptr_t) 15);
    gb = gb > tensor_end ? tensor_end : gb;  // (tensor base and size are multiples of 16: host-checked)
    seg[s].g = ga;
    seg[s].bytes = hi >= lo ? (uint32_t) (gb - ga) : 0u;
    seg[s].soff = soff;
    seg[s].n = n;
    seg[s].iy_lo = lo;
    seg[s].delta = (uint32_t) (g0 - ga);
    soff += seg[s].bytes;
  }
  return nseg;
}

__device__ __forceinline__ void run9_load_smem(uint32_t saddr, uint32_t (&w)[3]) {
  const uint32_t base = saddr & ~3u;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[0]) : "r"(base));
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[1]) : "r"(base + 4));
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[2]) : "r"(base + 8));
}

// Global-memory variant (used when the tensor cannot be bulk-copied: unaligned base / size, or tiny images)
__device__ __forceinline__ void load_a_conv_run9(const IgemmParams& p, const Item& it, uint32_t a_stage, int ltid) {
  constexpr int NB = 4;  // pixels (sub-tiles) in flight per thread
  const uint32_t fill = (uint32_t) p.izp * 0x01010101u;
  // pixel of this thread in sub-tile 0 by division (M < 2^31, host-checked); the following sub-tiles are 128 pixels
  // further along the same NHW order, reached by stepping — two divisions per item instead of two per pixel
  uint32_t cn;
  int coy, cox;
  {
    const uint32_t mu = (uint32_t) (it.m0 + ltid);
    const uint32_t t = mu / (uint32_t) p.out_w;
    cox = (int) (mu - t * (uint32_t) p.out_w);
    cn = t / (uint32_t) p.out_h;
    coy = (int) (t - cn * (uint32_t) p.out_h);
  }
  for (int j0 = 0; j0 < it.mt_eff; j0 += NB) {
    uint32_t w[NB][3][3];
    // per run: bits 0-1 = byte alignment of the source, bits 2-3 = state (0: in bounds, 1: all padding, 2: edge)
    uint32_t info[NB] = {};
#pragma unroll
    for (int jj = 0; jj < NB; jj++) {
      const long long m = it.m0 + (long long) (j0 + jj) * kTileM + ltid;
      const bool live = (j0 + jj) < it.mt_eff && m < p.M;
      const uint32_t n = live ? cn : 0u;
      const int oy = live ? coy : 0, ox = live ? cox : 0;
      cox += kTileM;  // advance to the same row of the next sub-tile
      while (cox >= p.out_w) {
        cox -= p.out_w;
        if (++coy == p.out_h) coy = 0, cn++;
      }
      const int iy0 = oy * p.stride_h - p.pad_top, ix0 = ox * p.stride_w - p.pad_left;
      const bool full = ix0 >= 0 && ix0 + p.kw <= p.in_w;
      const uint8_t* img = p.in + ((size_t) n * p.in_h * p.in_w + (full ? ix0 : 0)) * 3;
#pragma unroll
      for (int ky = 0; ky < 3; ky++) {
        const int iy = iy0 + ky * p.dil_h;
        const bool rowok = (unsigned) iy < (unsigned) p.in_h;
        const uint32_t st = (!live || !rowok) ? 1u : (full ? 0u : 2u);
        const uint8_t* src = img + (size_t) (rowok ? iy : 0) * p.in_w * 3;
        info[jj] |= (((uint32_t) reinterpret_cast<uintptr_t>(src) & 3u) | (st << 2)) << (4 * ky);
        if (st == 0) {
          run9_load(src, w[jj][ky]);
        } else {
          w[jj][ky][0] = w[jj][ky][1] = w[jj][ky][2] = fill;
        }
      }
    }
#pragma unroll
    for (int jj = 0; jj < NB; jj++) {
      if (j0 + jj < it.mt_eff) {
        uint32_t r[3][3];
#pragma unroll
        for (int ky = 0; ky < 3; ky++) {
          const uint32_t inf = info[jj] >> (4 * ky);
          const uint32_t st = (inf >> 2) & 3u;
          if (st == 0) {
            const uint32_t sh = (inf & 3u) * 8;
            r[ky][0] = __funnelshift_r(w[jj][ky][0], w[jj][ky][1], sh);
            r[ky][1] = __funnelshift_r(w[jj][ky][1], w[jj][ky][2], sh);
            r[ky][2] = (w[jj][ky][2] >> sh) & 0xFFu;
          } else if (st == 1) {
            r[ky][0] = fill, r[ky][1] = fill, r[ky][2] = fill & 0xFFu;
          } else {
            // left/right edge: assemble the run byte by byte (rare: one pixel per image row)
            const uint32_t mu = (uint32_t) (it.m0 + (long long) (j0 + jj) * kTileM + ltid);
            const uint32_t t = mu / (uint32_t) p.out_w;
            const int ox = (int) (mu - t * (uint32_t) p.out_w);
            const uint32_t n = t / (uint32_t) p.out_h;
            const int oy = (int) (t - n * (uint32_t) p.out_h);
            const int ix0 = ox * p.stride_w - p.pad_left;
            const int iy = oy * p.stride_h - p.pad_top + ky * p.dil_h;
            const uint8_t* rowp = p.in + ((size_t) n * p.in_h + iy) * p.in_w * 3;
            uint32_t b0 = 0, b1 = 0, b2 = 0;
            for (int i = 0; i < kRun; i++) {
              const int ix = ix0 + i / 3;
              const uint32_t v = ((unsigned) ix < (unsigned) p.in_w) ? (uint32_t) __ldg(rowp + (size_t) ix * 3 + i % 3)
                                                                      : (uint32_t) p.izp;
              if (i < 4) b0 |= v << (8 * i);
              else if (i < 8) b1 |= v << (8 * (i - 4));
              else b2 |= v;
            }
            r[ky][0] = b0, r[ky][1] = b1, r[ky][2] = b2;
          }
        }
        // K row: bytes [0,9) = ky 0, [9,18) = ky 1, [18,27) = ky 2, [27,32) = padding (zero weights)
        const uint32_t k0 = r[0][0], k1 = r[0][1];
        const uint32_t k2 = r[0][2] | (r[1][0] << 8);
        const uint32_t k3 = __funnelshift_r(r[1][0], r[1][1], 24);
        const uint32_t k4 = (r[1][1] >> 24) | (r[1][2] << 8) | (r[2][0] << 16);
        const uint32_t k5 = __funnelshift_r(r[2][0], r[2][1], 16);
        const uint32_t k6 = (r[2][1] >> 16) | (r[2][2] << 16);
        const uint32_t d = a_stage + (uint32_t) ((j0 + jj) * p.skc) * kChunkBytes + (uint32_t) ltid * 16;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(k0), "r"(k1), "r"(k2), "r"(k3) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + kChunkBytes), "r"(k4), "r"(k5), "r"(k6), "r"(0u)
                     : "memory");
      }
    }
  }
}

// The same K rows built from the raw-row staging buffer, branch-free.  Every run is read as 12 aligned bytes at its
// nominal address even when it starts left of the row or ends right of it (the neighbouring bytes are readable shared
// memory); the pixels of the run that are padding are then overwritten with the input zero point through byte masks,
// and a run whose whole row is padding is replaced likewise.  (The global-memory variant above needs an explicit
// byte-wise edge path for that; with a warp of consecutive pixels almost every third warp took it.)
__device__ __forceinline__ void load_a_conv_run9_raw(const IgemmParams& p, const Item& it, uint32_t a_stage, int ltid, int raw_n0,
                                                     uint32_t raw_b0, uint32_t raw_b1, uint32_t raw_safe) {
  const uint32_t fill = (uint32_t) p.izp * 0x01010101u;
  uint32_t cn;
  int coy, cox;
  {
    const uint32_t mu = (uint32_t) (it.m0 + ltid);  // M < 2^31 (host-checked)
    const uint32_t t = mu / (uint32_t) p.out_w;
    cox = (int) (mu - t * (uint32_t) p.out_w);
    cn = t / (uint32_t) p.out_h;
    coy = (int) (t - cn * (uint32_t) p.out_h);
  }
  const int row_pitch = p.in_w * 3;
#pragma unroll 2
  for (int j = 0; j < it.mt_eff; j++) {
    const bool live = it.m0 + (long long) j * kTileM + ltid < p.M;
    const int ix0 = cox * p.stride_w - p.pad_left, iy0 = coy * p.stride_h - p.pad_top;
    // byte masks of the padded pixels of a run: pixel i (bytes 3i..3i+2) is padding if ix0 + i is outside [0, in_w)
    const int lo = ix0 < 0 ? -ix0 : 0, hi = ix0 + 3 - p.in_w > 0 ? ix0 + 3 - p.in_w : 0;
    uint32_t m0 = lo == 0 ? 0u : (lo == 1 ? 0x00FFFFFFu : 0xFFFFFFFFu);
    uint32_t m1 = lo >= 2 ? (lo == 2 ? 0x0000FFFFu : 0xFFFFFFFFu) : 0u;
    uint32_t m2 = lo >= 3 ? 0xFFu : 0u;
    if (hi >= 1) m1 |= 0xFFFF0000u, m2 |= 0xFFu;
    if (hi >= 2) m0 |= 0xFF000000u, m1 = 0xFFFFFFFFu;
    if (hi >= 3) m0 = 0xFFFFFFFFu;
    const uint32_t base = ((int) cn == raw_n0 ? raw_b0 : raw_b1) + (uint32_t) (iy0 * row_pitch + ix0 * 3);
    uint32_t r[3][3];
    // (A warp-uniform "all nine taps inside the image" fast path without the masks was tried in round 2 and measured 4 %
    // slower than this branch-free form: the branch keeps the compiler from overlapping the loads of consecutive pixels.)
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
      const int iy = iy0 + ky * p.dil_h;
      const bool rowok = live && (unsigned) iy < (unsigned) p.in_h;
      const uint32_t a = base + (uint32_t) (ky * p.dil_h * row_pitch);
      uint32_t w[3];
      run9_load_smem(rowok ? a : raw_safe, w);  // (any readable address when the row is padding)
      const uint32_t sh = (a & 3u) * 8;
      const uint32_t r0 = __funnelshift_r(w[0], w[1], sh), r1 = __funnelshift_r(w[1], w[2], sh), r2 = w[2] >> sh;
      const uint32_t k0 = rowok ? m0 : 0xFFFFFFFFu, k1 = rowok ? m1 : 0xFFFFFFFFu, k2 = rowok ? m2 : 0xFFu;
      r[ky][0] = (r0 & ~k0) | (fill & k0);
      r[ky][1] = (r1 & ~k1) | (fill & k1);
      r[ky][2] = ((r2 & ~k2) | (fill & k2)) & 0xFFu;
    }
    // K row: bytes [0,9) = ky 0, [9,18) = ky 1, [18,27) = ky 2, [27,32) = padding (zero weights)
    const uint32_t k0 = r[0][0], k1 = r[0][1];
    const uint32_t k2 = r[0][2] | (r[1][0] << 8);
    const uint32_t k3 = __funnelshift_r(r[1][0], r[1][1], 24);
    const uint32_t k4 = (r[1][1] >> 24) | (r[1][2] << 8) | (r[2][0] << 16);
    const uint32_t k5 = __funnelshift_r(r[2][0], r[2][1], 16);
    const uint32_t k6 = (r[2][1] >> 16) | (r[2][2] << 16);
    const uint32_t d = a_stage + (uint32_t) (j * p.skc) * kChunkBytes + (uint32_t) ltid * 16;
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(k0), "r"(k1), "r"(k2), "r"(k3) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + kChunkBytes), "r"(k4), "r"(k5), "r"(k6), "r"(0u) : "memory");
    cox += kTileM;  // the same row of the next sub-tile is 128 pixels further along the NHW order
    while (cox >= p.out_w) {
      cox -= p.out_w;
      if (++coy == p.out_h) coy = 0, cn++;
    }
  }
}

// 1x1 / fully-connected through the TMA: the activation matrix is described once (host side) as a 3-D tensor
// {16 bytes of K, M rows, K/16 chunks}; ONE instruction then lands a [chunks][128 rows][16 B] box — exactly the
// no-swizzle UMMA operand image of a sub-tile — and signals the stage barrier with its byte count.  No per-lane
// address arithmetic, rows beyond M and chunks beyond K are zero-filled by the hardware.
constexpr int kVecTma = 32;  // value of the VEC template parameter that selects this loader
constexpr int kVecRaw9 = 2;  // 3x3x3 run loader fed from bulk-copied raw rows (load_a_conv_run9<true>)

__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* tmap, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          dst_smem),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void copy_bytes16(uint32_t dst, const uint8_t* src, int bytes, int ltid) {
  for (int o = ltid * 16; o < bytes; o += kLoadThreads * 16) cp_async<16>(dst + o, src + o);
}

// ------------------------------------------------------------------------------------------------
// epilogue
// ------------------------------------------------------------------------------------------------
// Slow path of the direct store: fewer than 16 valid bytes, or a destination that is not 16-byte aligned.
__device__ __noinline__ void store_row_partial(uint8_t* dst, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int valid,
                                               int vec) {
  const uint32_t w[4] = {w0, w1, w2, w3};
  int i = 0;
  if (vec >= 8) {
    for (; i + 8 <= valid; i += 8) *reinterpret_cast<uint2*>(dst + i) = make_uint2(w[i >> 2], w[(i >> 2) + 1]);
  }
  if (vec >= 4) {
    for (; i + 4 <= valid; i += 4) *reinterpret_cast<uint32_t*>(dst + i) = w[i >> 2];
  }
  for (; i < valid; i++) dst[i] = (uint8_t) (w[i >> 2] >> (8 * (i & 3)));
}

struct EpiCtx {
  uint32_t tlane;      // TMEM address of this warp's lane quarter, column 0 of the accumulator stage
  uint32_t bias_base;  // smem address of this (group, n_tile)'s folded biases ("ones" mode)
  uint32_t staging;    // smem staging buffer of the epilogue pair (bulk mode)
  uint8_t* obase;      // out + g*goc + nt*n_tile
  long long item;
  int row;             // row inside a sub-tile == TMEM lane
  int n_valid;         // valid output channels of this n-tile
  bool bulk;           // this item's output goes through the staging buffer and one bulk (1-D TMA) store
  // hand-over barriers, touched as late / as early as the data dependences allow (both waits used to sit at the top
  // of the item and showed up as the two largest stall sites of the kernel):
  uint32_t out_free_bar, out_free_parity;  // waited on right before the warp's FIRST staging write of the item
  uint32_t tmem_empty_bar;                 // arrived on right after the warp's LAST TMEM read of the item
};

// NB output bytes (NB/4 packed words, NB = 16 or 32) of row m, columns [c0, c0+NB) of the n-tile.
// bulk: into the dense smem image of the item's output rows (pitch goc, goc % 4 == 0), later written by ONE
// cp.async.bulk per item — the TMA engine produces full-width global writes, which strided per-thread stores do not.
// FAST: the item is known to be a full bulk item of full 16-column groups -> unconditional 16-byte staging stores
template <int NB, bool FAST>
__device__ __forceinline__ void emit(const IgemmParams& p, const Item& it, const EpiCtx& e, int j, int c0, const uint32_t* w) {
  if constexpr (FAST) {
    const uint32_t s = e.staging + (uint32_t) (j * kTileM + e.row) * p.goc + c0;
#pragma unroll
    for (int h = 0; h < NB / 16; h++)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s + 16 * h), "r"(w[4 * h]), "r"(w[4 * h + 1]), "r"(w[4 * h + 2]),
                   "r"(w[4 * h + 3])
                   : "memory");
    return;
  }
  const int valid = e.n_valid - c0;
  if (valid <= 0) return;
  if (e.bulk) {
    const uint32_t s = e.staging + (uint32_t) (j * kTileM + e.row) * p.goc + c0;
#pragma unroll
    for (int h = 0; h < NB / 16; h++) {
      const int v = valid - 16 * h;
      if (v >= 16 && (p.goc & 15) == 0) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s + 16 * h), "r"(w[4 * h]), "r"(w[4 * h + 1]),
                     "r"(w[4 * h + 2]), "r"(w[4 * h + 3])
                     : "memory");
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (4 * k < v) asm volatile("st.shared.b32 [%0], %1;" ::"r"(s + 16 * h + 4 * k), "r"(w[4 * h + k]) : "memory");
      }
    }
    return;
  }
  const long long m = it.m0 + (long long) j * kTileM + e.row;
  if (m >= p.M) return;
  uint8_t* dst = e.obase + (size_t) m * p.out_stride + c0;
  if constexpr (NB == 32) {
    if (valid >= 32 && p.out_vec >= 32) {
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                   "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                   : "memory");
      return;
    }
  }
#pragma unroll
  for (int h = 0; h < NB / 16; h++) {
    const int v = valid - 16 * h;
    if (v <= 0) break;
    if (v >= 16 && p.out_vec >= 16) {
      *reinterpret_cast<uint4*>(dst + 16 * h) = make_uint4(w[4 * h], w[4 * h + 1], w[4 * h + 2], w[4 * h + 3]);
    } else {
      store_row_partial(dst + 16 * h, w[4 * h], w[4 * h + 1], w[4 * h + 2], w[4 * h + 3], v < 16 ? v : 16, p.out_vec);
    }
  }
}

// W (16 or 32) accumulator columns of one row: TMEM -> registers -> requantise -> pack -> store.
// FOLDED: the accumulator already contains bias and zero-point correction (extra UMMAs); otherwise
// ("ones" mode) the folded bias comes from smem and -kzp*rowsum from accumulator column n_tile.
template <int RQ, int W, bool FOLDED, bool FAST>
__device__ __forceinline__ void epilogue_cols(const IgemmParams& p, const Item& it, const EpiCtx& e, int j, int c0, bool first,
                                              bool last) {
  int32_t v[W];
  const uint32_t taddr = e.tlane + j * p.n_mma + c0;
  if constexpr (W == 32) {
    tmem_ld32(taddr, v);
  } else {
    tmem_ld16(taddr, v);
  }
  int32_t rowsum = 0;
  if constexpr (!FOLDED) {
    if (p.has_corr) tmem_ld1(e.tlane + j * p.n_mma + p.n_tile, rowsum);
  }
  tmem_ld_wait();
  if (last) {  // the accumulator stage is drained as far as this warp is concerned: let the UMMA warp refill it now
    tc_fence_before_sync();
    mbar_arrive(e.tmem_empty_bar);
  }
  if (!FAST && p.dbg_acc != nullptr) {  // bring-up aid; v[] is only indexed with constants, so it stays in registers
    int32_t* d = p.dbg_acc + (((size_t) e.item * p.mt + j) * kTileM + e.row) * p.n_mma;
#pragma unroll
    for (int i = 0; i < W; i++) d[c0 + i] = v[i];
    if (!FOLDED && c0 == 0 && p.has_corr) d[p.n_tile] = rowsum;
  }
  constexpr bool kU = (RQ == 5 || RQ == 6);  // "U" requantisation takes n XOR 2^31, i.e. n + 2^31 (mod 2^32)
  if constexpr (!FOLDED) {
    const int32_t corr = (int32_t) ((uint32_t) (-p.kzp * rowsum) + (kU ? 0x80000000u : 0u));  // offset rides on the bias add
#pragma unroll
    for (int t = 0; t < W / 4; t++) {
      int4 b;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                   : "r"(e.bias_base + (uint32_t) (c0 + 4 * t) * 4));
      v[4 * t + 0] += b.x + corr;
      v[4 * t + 1] += b.y + corr;
      v[4 * t + 2] += b.z + corr;
      v[4 * t + 3] += b.w + corr;
    }
  }
  uint32_t w[W / 4];
  if constexpr (kU) {
    const uint32_t m2 = p.rq.u_m2, flip = FOLDED ? 0x80000000u : 0u;
    const uint64_t k2 = p.rq.u_k2;
    const int32_t sm = p.rq.u_sm;
    auto rq1 = [&](int32_t n) -> int32_t {
      int32_t y = q8_requant_u_unclamped((uint32_t) n ^ flip, m2, k2, sm);
      if constexpr (RQ == 6) y = min(max(y, p.rq.qmin), p.rq.qmax);
      return y;
    };
#pragma unroll
    for (int t = 0; t < W / 4; t++) w[t] = pack_sat_u8x4(rq1(v[4 * t]), rq1(v[4 * t + 1]), rq1(v[4 * t + 2]), rq1(v[4 * t + 3]));
  } else {
#pragma unroll
    for (int t = 0; t < W / 4; t++) w[t] = requant_pack4_generic(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3], p.rq);
  }
  if (first && p.out_mode == 1) mbar_wait(e.out_free_bar, e.out_free_parity);  // the previous bulk store has left staging
  emit<W, FAST>(p, it, e, j, c0, w);
}

// One epilogue warp = lane quarter q (warp % 4) of its pair's accumulator stage; the two warps of a pair that share
// a quarter take alternate (sub-tile, column-block) units of the item.
template <int RQ, bool FOLDED, bool FAST>
__device__ __forceinline__ void epilogue_item(const IgemmParams& p, const Item& it, const EpiCtx& e, int half) {
  constexpr int W = FOLDED ? 32 : 16;
  const int full = p.n_tile / W;            // full-width units per sub-tile
  const int per_sub = full + ((p.n_tile % W) ? 1 : 0);
  const int units = it.mt_eff * per_sub;
  int j = half / per_sub, c = half - j * per_sub;
  if (half >= units) {  // nothing to read for this warp (tail item): release the stage right away
    tc_fence_before_sync();
    mbar_arrive(e.tmem_empty_bar);
  }
  for (int u = half; u < units; u += 2) {
    const bool first = u == half, last = u + 2 >= units;
    if (c < full) {
      epilogue_cols<RQ, W, FOLDED, FAST>(p, it, e, j, c * W, first, last);
    } else {
      epilogue_cols<RQ, 16, FOLDED, FAST>(p, it, e, j, c * W, first, last);  // 16-column remainder (FOLDED, n_tile % 32 == 16)
    }
    c += 2;
    while (c >= per_sub) {
      c -= per_sub;
      ++j;
    }
  }
}

template <int RQ>
__device__ __forceinline__ void epilogue_dispatch(const IgemmParams& p, const Item& it, const EpiCtx& e, int half) {
  // fast variant (specialised requantisation forms only): full bulk item, every 16-column group complete, no debug dump
  const bool fast = (RQ == 5 || RQ == 6) && e.bulk && p.dbg_acc == nullptr && (p.goc & 15) == 0 && e.n_valid == p.n_tile;
  if (p.folded) {
    if (fast) {
      epilogue_item<RQ, true, (RQ == 5 || RQ == 6)>(p, it, e, half);
    } else {
      epilogue_item<RQ, true, false>(p, it, e, half);
    }
  } else {
    if (fast) {
      epilogue_item<RQ, false, (RQ == 5 || RQ == 6)>(p, it, e, half);
    } else {
      epilogue_item<RQ, false, false>(p, it, e, half);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// panel epilogue (out_mode 2)
// ------------------------------------------------------------------------------------------------
// Measured on the round-1 kernel (ncu source view, profiles/r1h_igemm_first3): the epilogue warps are busy ~90 % of
// the time and only 152 of the ~260 instructions they execute per 32-column unit are the requantisation itself; the
// rest was per-unit scaffolding (item/tail/alignment cases, address arithmetic, re-convergence), and the two IMAD.HI
// per value keep the half-rate "heavy" FMA pipe busier (64 %) than any other unit.  This path has no cases at all:
//   * every item — full, tail, ragged N, strided output — leaves through swizzled smem panels and 2-D tensor stores
//     (the TMA clips rows >= M and columns >= N), so the unit loop is  load - requantise - two 16-byte stores;
//   * panel pitches are 128/64/32/16 bytes with the matching TMA swizzle: the 8 lanes of a store phase always hit 8
//     different 16-byte bank groups (the dense N-byte pitch gave 2- to 8-way conflicts for N = 96 ... 1280);
//   * the final arithmetic shift alternates between IMAD.HI (heavy FMA pipe, 4 cycles per warp) and SHF (ALU pipe,
//     2 cycles), which balances the two pipes at ~6 cycles per 32 values instead of 8 on the FMA pipe alone.
template <int RQ, bool FOLDED>
__device__ __forceinline__ void epi2_requant16(const IgemmParams& p, const int32_t* v, uint32_t* w) {
  const uint32_t m2 = p.rq.u_m2, flip = FOLDED ? 0x80000000u : 0u;  // ("ones" mode: the offset rides on the bias add)
  const uint64_t k2 = p.rq.u_k2;
  const int32_t sm = p.rq.u_sm, sh = p.rq.shift;
  auto rq = [&](int32_t n, bool mul) -> int32_t {
    const uint32_t nu = (uint32_t) n ^ flip;
    const uint32_t hi = (uint32_t) (((uint64_t) nu * m2 + k2) >> 32);
    const int32_t t = (int32_t) (hi + (nu >> 31));
    int32_t y = mul ? __mulhi(t, sm) : (t >> sh);  // same value (sm = 2^(32 - sh)), different pipe
    if constexpr (RQ == 6) y = min(max(y, p.rq.qmin), p.rq.qmax);
    return y;
  };
#pragma unroll
  for (int t = 0; t < 4; t++)
    w[t] = pack_sat_u8x4(rq(v[4 * t], true), rq(v[4 * t + 1], false), rq(v[4 * t + 2], true), rq(v[4 * t + 3], false));
}

// One unit: W accumulator columns of the warp's 32 rows.  a0 / a1 = smem addresses of the unit's first / second 16-byte
// chunk in this lane's staging row.
// How an epilogue warp reports "I have read all I need from this accumulator stage":
//   local : every lane arrives on the CTA's own tmem_empty barrier (count = threads of the epilogue group)
//   remote: CTA pairs (q8_gemm2sm_kernel) — the UMMA-issuing thread lives in the pair's leader CTA, so ONE lane per warp
//           arrives on the leader's barrier through its shared::cluster address (count = warps of both CTAs)
template <bool REMOTE>
__device__ __forceinline__ void release_tmem_stage(uint32_t tmem_empty_bar) {
  tc_fence_before_sync();
  if constexpr (REMOTE) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive_cluster(tmem_empty_bar);
  } else {
    mbar_arrive(tmem_empty_bar);
  }
}

template <int RQ, int W, bool FOLDED, bool REMOTE = false>
__device__ __forceinline__ void epi2_unit(const IgemmParams& p, uint32_t taddr, uint32_t rs_taddr, uint32_t bias_addr, uint32_t a0,
                                          uint32_t a1, bool first, bool last, uint32_t tmem_empty_bar, uint32_t out_free_bar,
                                          uint32_t out_free_parity) {
  int32_t v[W];
  if constexpr (W == 32) {
    tmem_ld32(taddr, v);
  } else {
    tmem_ld16(taddr, v);
  }
  int32_t rowsum = 0;
  if constexpr (!FOLDED) tmem_ld1(rs_taddr, rowsum);
  tmem_ld_wait();
  if (last) release_tmem_stage<REMOTE>(tmem_empty_bar);  // all read: the UMMA warps may refill the accumulator stage
  if constexpr (!FOLDED) {
    const int32_t corr = (int32_t) ((uint32_t) (-p.kzp * rowsum) + 0x80000000u);  // "U" offset rides on the bias add
#pragma unroll
    for (int t = 0; t < W / 4; t++) {
      int4 b;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(bias_addr + 16 * t));
      v[4 * t + 0] += b.x + corr;
      v[4 * t + 1] += b.y + corr;
      v[4 * t + 2] += b.z + corr;
      v[4 * t + 3] += b.w + corr;
    }
  }
  uint32_t w[W / 4];
  epi2_requant16<RQ, FOLDED>(p, v, w);
  if constexpr (W == 32) epi2_requant16<RQ, FOLDED>(p, v + 16, w + 4);
  if (first) mbar_wait(out_free_bar, out_free_parity);  // the tensor stores of the pair's previous item have read staging
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
  if constexpr (W == 32)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}

// All units of one item for warp (quarter q, half): the two warps of a quarter take alternate units of the item's
// (sub-tile, column block) sequence.  Everything that does not change inside the loop is an argument.
template <int RQ, bool FOLDED, bool REMOTE = false>
__device__ __forceinline__ void epi2_item(const IgemmParams& p, int mt_eff, uint32_t tlane, uint32_t bias_base, uint32_t staging,
                                          uint32_t row, int half, uint32_t tmem_empty_bar, uint32_t out_free_bar,
                                          uint32_t out_free_parity, int n_tile) {
  constexpr int W = FOLDED ? 32 : 16;
  const int full = n_tile / W;
  const int per_sub = full + ((n_tile % W) ? 1 : 0);  // (W == 32: a 16-column remainder unit when n_tile % 32 == 16)
  const int units = mt_eff * per_sub;
  if (half >= units) {  // a single-unit item: nothing for the second warp of the quarter
    release_tmem_stage<REMOTE>(tmem_empty_bar);
    return;
  }
  int c = half, j = 0;
  if (c >= per_sub) c -= per_sub, j = 1;  // (per_sub == 1)
  uint32_t jrow = (uint32_t) j * kTileM + row;
  uint32_t tsub = tlane + (uint32_t) (j * p.n_mma);
#pragma unroll 1
  for (int u = half; u < units; u += 2) {
    const bool first = u == half, last = u + 2 >= units;
    // staging address of the unit's first chunk: panel base + row pitch, chunk bits swizzled like the panel's TMA mode
    const uint2 ent = p.e2_unit[c];
    const uint32_t pitch = ent.y & 0xFFu, lsh = (ent.y >> 8) & 0xFFu, mask = ent.y >> 16;
    const uint32_t a0 = (staging + ent.x + jrow * pitch) ^ ((row << lsh) & mask);
    // second chunk of a 32-column unit: in a swizzled panel the unit starts on a 32-byte boundary of a 32-byte-multiple
    // pitch, so the neighbouring chunk differs in address bit 4 only; the dense image has pitch N with N / 16 odd
    const uint32_t a1 = p.e2_dense ? a0 + 16u : a0 ^ 16u;
    const uint32_t taddr = tsub + (uint32_t) (c * W);
    if (c < full) {
      epi2_unit<RQ, W, FOLDED, REMOTE>(p, taddr, tsub + n_tile, bias_base + (uint32_t) (c * W) * 4, a0, a1, first, last, tmem_empty_bar,
                               out_free_bar, out_free_parity);
    } else {
      epi2_unit<RQ, 16, FOLDED, REMOTE>(p, taddr, tsub + n_tile, bias_base + (uint32_t) (c * W) * 4, a0, a1, first, last, tmem_empty_bar,
                                out_free_bar, out_free_parity);
    }
    c += 2;
    if (c >= per_sub) {
      c -= per_sub;
      jrow += kTileM;
      tsub += (uint32_t) p.n_mma;
      if (c >= per_sub) {  // (per_sub == 1: both warps advance one sub-tile per unit... two per step)
        c -= per_sub;
        jrow += kTileM;
        tsub += (uint32_t) p.n_mma;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int MODE, int VEC>
__global__ void __launch_bounds__(kThreads, 1)
    q8_igemm_kernel(const __grid_constant__ IgemmParams p, const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ IgemmStoreMaps smaps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ SmemCtl ctl;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_smem = smem_base + p.smem_b_off;
  const uint32_t bias_smem = smem_base + p.smem_bias_off;
  const uint32_t a_smem = smem_base + p.smem_a_off;

  if (tid == 0) {
    for (int s = 0; s < p.num_stages; s++) {
      // arrivals per stage: the 128 cp.async loaders, or one expect_tx arrival (TMA) plus the loaders when they stream B
      // arrivals per stage: one expect_tx arrival when the TMA thread moves everything (A tiles and, when they stream,
      // the weights), else the 128 cp.async loaders
      mbar_init(smem_u32(&ctl.full[s]), VEC == kVecTma ? 1 : kLoadThreads);
      mbar_init(smem_u32(&ctl.empty[s]), kMmaWarps);  // one tcgen05.commit per issuing warp
    }
    for (int s = 0; s < kMaxAccStages; s++) {
      mbar_init(smem_u32(&ctl.tmem_full[s]), kMmaWarps);
      mbar_init(smem_u32(&ctl.tmem_empty[s]), kEpiPairThreads);
    }
    mbar_init(smem_u32(&ctl.b_full), kLoadThreads);
    for (int s = 0; s < 2; s++) {
      mbar_init(smem_u32(&ctl.out_full[2 * s]), kEpiPairThreads);
      mbar_init(smem_u32(&ctl.out_free[2 * s]), 1);
      mbar_init(smem_u32(&ctl.out_full[2 * s + 1]), kEpiPairThreads);
      mbar_init(smem_u32(&ctl.out_free[2 * s + 1]), 1);
    }
    for (int s = 0; s < kMaxRawBufs; s++) {
      mbar_init(smem_u32(&ctl.raw_full[s]), 1);
      mbar_init(smem_u32(&ctl.raw_empty[s]), kLoadThreads);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc<kTmemCols>(smem_u32(&ctl.tmem_base));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl.tmem_base;

  const long long first = blockIdx.x, step = gridDim.x;

  if (warp >= kLoadWarp0 && warp < kStoreWarp) {
    // ===================================== loaders =====================================
    const int ltid = tid - kLoadWarp0 * 32;
    {
      // one-time: folded biases (always) and the packed weights (when they fit) become smem-resident
      copy_bytes16(bias_smem, reinterpret_cast<const uint8_t*>(p.bias), p.bias_count * 4, ltid);
      if (p.b_resident) copy_bytes16(b_smem, p.wpack, p.groups * p.n_tiles * p.blk_chunks * p.n_mma * 16, ltid);
      if (p.folded) {
        // constant A operand of the bias UMMAs: every row = [255 x 31, 1]  (chunk 0 = k 0..15, chunk 1 = k 16..31)
        const uint32_t ac = smem_base + p.smem_aconst_off + (uint32_t) ltid * 16;
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ac), "r"(0xFFFFFFFFu) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %2};" ::"r"(ac + kChunkBytes), "r"(0xFFFFFFFFu), "r"(0x01FFFFFFu)
                     : "memory");
        fence_proxy_async_smem();
      }
      cp_async_mbar_arrive_noinc(smem_u32(&ctl.b_full));
    }
    int stage = 0;
    uint32_t phase = 0;
    // raw staging ring: buffer rb (use parity rpar) holds the rows of the current item; they are requested by a thread of
    // the store warp (see there), which keeps up to raw_bufs items of rows in flight
    [[maybe_unused]] int rb = 0;
    [[maybe_unused]] uint32_t rpar = 0;
    [[maybe_unused]] const uint32_t raw0 = smem_base + (uint32_t) p.smem_raw_off;
    for (long long item = first; item < p.total_items; item += step) {
      const Item it = decode_item(p, item);
      for (int ks = 0; ks < p.k_stages; ks++) {
        if constexpr (VEC == kVecTma) {
          if (ltid != 0) continue;  // one thread drives the TMA (activations as tensor boxes, streamed weights as bulk copies)
        }
        mbar_wait_relaxed(smem_u32(&ctl.empty[stage]), phase ^ 1, 32);
        const uint32_t a_stage = a_smem + stage * p.stage_bytes;
        if constexpr (VEC == kVecTma) {
          if (ltid == 0) {
            const uint32_t bar = smem_u32(&ctl.full[stage]);
            // streamed weights: the stage's K chunks of this (group, n-tile) block are ONE contiguous run of the packed
            // blob, moved by bulk copies on the same barrier (round 1 used 128 threads x cp.async for it: the large-GEMM
            // and N >= 320 layers spent their time in those loader warps)
            int cs = p.nkc - ks * p.skc;
            cs = cs < p.skc ? cs : p.skc;
            const uint32_t b_bytes = p.b_resident ? 0u : (uint32_t) (cs * p.n_mma) * 16u;
            mbar_arrive_expect_tx(bar, (uint32_t) (it.mt_eff * p.skc) * kChunkBytes + b_bytes);
            for (int j = 0; j < it.mt_eff; j++) {
              const uint32_t dst = a_stage + (uint32_t) (j * p.skc) * kChunkBytes;
              const int row0 = (int) (it.m0 + (long long) j * kTileM);
              tma_load_3d(dst, &tmap_a, 0, row0, p.a_sw32 ? (ks * p.skc) >> 1 : ks * p.skc, bar);
              if (p.a_sw32 == 2)  // K % 32 == 16: the last 16 bytes of K (+ a zero-filled partner chunk) behind the slabs
                tma_load_3d(dst + (uint32_t) p.a_tail_c * kChunkBytes, reinterpret_cast<const CUtensorMap*>(&smaps.m[4][0]), 0, row0,
                            p.a_tail_c, bar);
            }
            if (!p.b_resident) {
              const uint8_t* wsrc =
                  p.wpack + ((size_t) (it.g * p.n_tiles + it.nt) * p.blk_chunks + (size_t) ks * p.skc) * p.n_mma * 16;
              const uint32_t b_dst = a_stage + (uint32_t) (p.mt * p.skc) * kChunkBytes;
              for (uint32_t o = 0; o < b_bytes; o += 16384) bulk_g2s(b_dst + o, wsrc + o, b_bytes - o < 16384 ? b_bytes - o : 16384u, bar);
            }
          }
        } else if constexpr (MODE == kModeGemm) {
          load_a_gemm<VEC>(p, it, ks, a_stage, ltid);
        } else if constexpr (VEC == 0) {
          load_a_conv_run9(p, it, a_stage, ltid);  // K = 27 fits one stage
        } else if constexpr (VEC == kVecRaw9) {
          RawSeg sg[2];
          const int ns = raw_segments(p, it, sg);
          const uint32_t rbase = raw0 + (uint32_t) rb * (uint32_t) p.raw_cap;
          const uint32_t b0 = rbase + sg[0].soff + sg[0].delta - (uint32_t) (sg[0].iy_lo * p.in_w) * 3u;
          const uint32_t b1 = ns > 1 ? rbase + sg[1].soff + sg[1].delta - (uint32_t) (sg[1].iy_lo * p.in_w) * 3u : b0;
          mbar_wait(smem_u32(&ctl.raw_full[rb]), rpar);
          load_a_conv_run9_raw(p, it, a_stage, ltid, sg[0].n, b0, b1, rbase);
          mbar_arrive(smem_u32(&ctl.raw_empty[rb]));
          if (++rb == p.raw_bufs) rb = 0, rpar ^= 1;
        } else {
          load_a_conv<VEC>(p, it, ks, a_stage, ltid);
        }
        if (VEC != kVecTma && !p.b_resident) {
          int cs = p.nkc - ks * p.skc;
          cs = cs < p.skc ? cs : p.skc;
          const uint8_t* wsrc =
              p.wpack + ((size_t) (it.g * p.n_tiles + it.nt) * p.blk_chunks + (size_t) ks * p.skc) * p.n_mma * 16;
          copy_bytes16(a_stage + p.mt * p.skc * kChunkBytes, wsrc, cs * p.n_mma * 16, ltid);
        }
        if constexpr (VEC != kVecTma) {
          fence_proxy_async_smem();  // st.shared fills (padding taps / byte path) -> UMMA reads
          cp_async_mbar_arrive_noinc(smem_u32(&ctl.full[stage]));
        }
        if (++stage == p.num_stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    cp_async_wait_all();
  } else if (warp >= kMmaWarp && warp < kMmaWarp + kMmaWarps) {
    // ===================================== UMMA issue =====================================
    // The whole warp walks the loop CONVERGED and every operand below is warp-uniform by construction (kernel
    // parameters, blockIdx, loop counters, values broadcast with __shfl_sync), so the compiler keeps descriptors and
    // addresses in uniform registers and a tcgen05.mma costs its operand arithmetic plus ONE instruction; one elected
    // lane issues.  (With a single-lane branch around the loop nvcc had to rebuild every operand in a vector register
    // and move it across with an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall: ~30 instructions per UMMA.)
    {
      const int w = __shfl_sync(0xffffffffu, warp - kMmaWarp, 0);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_u = __shfl_sync(0xffffffffu, smem_base, 0);
      const uint32_t ctl_u = __shfl_sync(0xffffffffu, smem_u32(&ctl), 0);
      const uint32_t bar_full = ctl_u + (uint32_t) offsetof(SmemCtl, full), bar_empty = ctl_u + (uint32_t) offsetof(SmemCtl, empty);
      const uint32_t bar_tfull = ctl_u + (uint32_t) offsetof(SmemCtl, tmem_full);
      const uint32_t bar_tempty = ctl_u + (uint32_t) offsetof(SmemCtl, tmem_empty);
      const uint32_t b_smem_u = smem_u + p.smem_b_off, a_smem_u = smem_u + p.smem_a_off;
      const uint32_t idesc_main = umma_idesc_i8(kTileM, (uint32_t) p.n_mma, false, p.b_signed != 0);
      const uint32_t idesc_us = umma_idesc_i8(kTileM, (uint32_t) p.n_mma, false, true);  // u8 x s8
      const uint32_t b_lbo = (uint32_t) p.n_mma * 16;
      const uint32_t sub_bytes = (uint32_t) p.skc * kChunkBytes;
      const uint32_t a_const = smem_u + p.smem_aconst_off;
      // descriptor templates (strides, version); the 16-byte-granular start address is added per use — every operand
      // lies below 256 KB, so the 14-bit address field cannot carry into its neighbours
      const uint64_t a_tmpl = umma_desc_kmajor_noswizzle(0, kChunkBytes, 128);
      // activations: 16-byte chunks (no swizzle: two chunks = one K = 32 step, LBO apart) or 32-byte slabs written by the TMA
      // with the 32-byte swizzle (one slab = one K step; layout type 6, 8-row groups 256 bytes apart).  Slab s sits where
      // chunks 2s, 2s+1 sat, so the per-step byte offsets below are the same for both.
      const uint64_t a_main = p.a_sw32 ? (((uint64_t) 1 << 16) | ((uint64_t) (256u >> 4) << 32) | ((uint64_t) 1 << 46) | ((uint64_t) 6 << 61))
                                       : a_tmpl;
      const uint64_t b_tmpl = umma_desc_kmajor_noswizzle(0, b_lbo, 128);
      if (p.b_resident) {
        mbar_wait(ctl_u + (uint32_t) offsetof(SmemCtl, b_full), 0);
        fence_proxy_async_smem();
      }
      int stage = 0;
      uint32_t phase = 0;
      int as = -1;             // accumulator stage of the item: round-robin over the 2..4 TMEM stages, so the UMMAs
      uint32_t as_phase = 1;   // of the next items run while the epilogue still drains earlier ones
      for (long long item = first; item < p.total_items; item += step) {
        const Item it = decode_item(p, item);
        if (++as == p.acc_stages) as = 0;
        as_phase ^= (as == 0);
        mbar_wait_parked(bar_tempty + 8u * (uint32_t) as, as_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_u + as * p.acc_stride;
        // base of this (group, n_tile)'s resident block: [B1: nkc][B2 full: 2][B2 tail: 2][bias digits: 2 per step]
        const uint32_t blk = b_smem_u + (uint32_t) ((it.g * p.n_tiles + it.nt) * p.blk_chunks) * b_lbo;
        for (int ks = 0; ks < p.k_stages; ks++) {
          mbar_wait_parked(bar_full + 8u * (uint32_t) stage, phase);
          fence_proxy_async_smem();
          tc_fence_after_sync();
          const uint32_t a_stage = a_smem_u + stage * p.stage_bytes;
          int cs = p.nkc - ks * p.skc;
          cs = cs < p.skc ? cs : p.skc;
          const uint32_t b_base = p.b_resident ? blk + (uint32_t) (ks * p.skc) * b_lbo : a_stage + p.mt * sub_bytes;
          if (elect_one()) {
            // UMMAs that accumulate into the same TMEM columns serialise (each waits for the previous result), so
            // the sub-tiles are the INNER loop: consecutive instructions hit different accumulators and pipeline.
            if (p.folded && ks == 0) {
              // accumulator := folded bias  (A = [255 x31, 1] in every row, B = signed base-255 digits)
              for (int t = 0; t < p.bias_steps; t++) {
                const uint64_t ad = a_tmpl + (a_const >> 4);
                const uint64_t bd = b_tmpl + ((blk + (uint32_t) (p.nkc + 4 + 2 * t) * b_lbo) >> 4);
                for (int j = w; j < it.mt_eff; j += kMmaWarps) umma_i8(d_tmem + j * p.n_mma, ad, bd, idesc_us, t != 0 ? 1u : 0u);
              }
            }
            for (int c = 0; c < cs; c += 2) {
              const uint32_t acc = (p.folded || (ks | c) != 0) ? 1u : 0u;
              const uint64_t bd = b_tmpl + ((b_base + c * b_lbo) >> 4);
              const uint64_t ad0 = ((p.a_sw32 == 2 && c >= p.a_tail_c) ? a_tmpl : a_main) + ((a_stage + c * kChunkBytes) >> 4);
              const uint32_t sub16 = sub_bytes >> 4;
              for (int j = w; j < it.mt_eff; j += kMmaWarps)
                umma_i8(d_tmem + j * p.n_mma, ad0 + (uint32_t) j * sub16, bd, idesc_main, acc);
              if (p.has_b2) {
                // + (128 - kzp) * sum_k a[m][k]: the zero-point correction as a second UMMA on the same A tile
                const bool tail = p.k_tail_pad && (ks * p.skc + c + 2 == p.nkc);
                const uint64_t b2 = b_tmpl + ((blk + (uint32_t) (p.nkc + (tail ? 2 : 0)) * b_lbo) >> 4);
                for (int j = w; j < it.mt_eff; j += kMmaWarps)
                  umma_i8(d_tmem + j * p.n_mma, ad0 + (uint32_t) j * sub16, b2, idesc_us, 1u);
              }
            }
            umma_commit(bar_empty + 8u * (uint32_t) stage);
            if (ks == p.k_stages - 1) umma_commit(bar_tfull + 8u * (uint32_t) as);
          }
          __syncwarp();
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp < kEpiWarps) {
    // ===================================== epilogue =====================================
    // Pair 0 (warps 0-7) handles even work items in TMEM stage 0, pair 1 (warps 8-15) odd items in stage 1.
    // A pair's eight warps never synchronise with each other: they hand finished output tiles to the store
    // thread through two mbarriers.
    const int pair = warp >> 3;
    const int q = warp & 3, half = (warp >> 2) & 1;
    const int lane = tid & 31;
    // staging: one buffer per pair, or two used alternately (staging_bufs == 2, panel epilogue only) — with one, the pair's
    // first staging write of an item waits until the tensor stores of its previous item have READ the buffer, which
    // the round-2 profile of the 16 -> 96 expansion showed as 8 % of the epilogue warps' time
    const uint32_t staging = smem_base + p.smem_stage_off + pair * p.staging_bufs * p.staging_bytes;
    mbar_wait(smem_u32(&ctl.b_full), 0);  // biases are in smem ("ones" mode reads them)
    if (p.out_mode == 2) {
      // panel epilogue: one specialised loop per (requantisation form, mode), chosen once per launch
      const uint32_t row = (uint32_t) (q * 32 + lane);
      auto run = [&](auto rq_tag, auto folded_tag) {
        constexpr int RQ = decltype(rq_tag)::value;
        constexpr bool FOLDED = decltype(folded_tag)::value;
        int as = pair - 2;
        uint32_t as_phase = 0, k = ~0u;
        for (long long item = first + pair * step; item < p.total_items; item += 2 * step) {
          const Item it = decode_item(p, item);
          k++;
          as += 2;
          if (as >= p.acc_stages) {
            as -= p.acc_stages;
            as_phase ^= 1;
          }
          mbar_wait(smem_u32(&ctl.tmem_full[as]), as_phase);
          tc_fence_after_sync();
          // buffer and the use count of that buffer's barriers (k-th item of the pair)
          const uint32_t buf = p.staging_bufs == 2 ? (k & 1u) : 0u, use = p.staging_bufs == 2 ? (k >> 1) : k;
          epi2_item<RQ, FOLDED>(p, it.mt_eff, tmem_base + as * p.acc_stride + ((uint32_t) (q * 32) << 16),
                                bias_smem + (uint32_t) ((it.g * p.n_tiles + it.nt) * p.n_tile) * 4,
                                staging + buf * (uint32_t) p.staging_bytes, row, half, smem_u32(&ctl.tmem_empty[as]),
                                smem_u32(&ctl.out_free[2 * pair + buf]), (use & 1) ^ 1, p.n_tile);
          fence_proxy_async_smem();  // staging writes (generic proxy) -> tensor stores (async proxy)
          mbar_arrive(smem_u32(&ctl.out_full[2 * pair + buf]));
        }
      };
      using T5 = std::integral_constant<int, 5>;
      using T6 = std::integral_constant<int, 6>;
      if (p.folded) {
        if (p.rq_mode == 5) run(T5{}, std::true_type{}); else run(T6{}, std::true_type{});
      } else {
        if (p.rq_mode == 5) run(T5{}, std::false_type{}); else run(T6{}, std::false_type{});
      }
    } else {
    int as = pair - 2;        // accumulator stage = (local item index) mod acc_stages, tracked without divisions
    uint32_t as_phase = 0;    // = ((local item index) / acc_stages) & 1
    uint32_t k = ~0u;         // this pair's item counter
    for (long long item = first + pair * step; item < p.total_items; item += 2 * step) {
      const Item it = decode_item(p, item);
      k++;
      as += 2;
      if (as >= p.acc_stages) {
        as -= p.acc_stages;
        as_phase ^= 1;
      }
      const bool bulk = p.out_mode == 1 && (it.m0 + (long long) it.mt_eff * kTileM <= p.M);
      mbar_wait(smem_u32(&ctl.tmem_full[as]), as_phase);
      tc_fence_after_sync();
      EpiCtx e;
      e.tlane = tmem_base + as * p.acc_stride + ((uint32_t) (q * 32) << 16);
      e.bias_base = bias_smem + (uint32_t) ((it.g * p.n_tiles + it.nt) * p.n_tile) * 4;
      e.staging = staging;
      e.obase = p.out + (size_t) it.g * p.goc + (size_t) it.nt * p.n_tile;
      e.item = item;
      e.row = q * 32 + lane;
      e.n_valid = min(p.n_tile, p.goc - it.nt * p.n_tile);
      e.bulk = bulk;
      e.out_free_bar = smem_u32(&ctl.out_free[2 * pair]);
      e.out_free_parity = (k & 1) ^ 1;
      e.tmem_empty_bar = smem_u32(&ctl.tmem_empty[as]);
      // two specialised epilogues ("U" requantisation without / with clamp: every layer whose accumulators are bounded,
      // i.e. practically all) and one generic one; more instantiations only bloat the kernel (215 KB of code before)
      switch (p.rq_mode) {
        case 5: epilogue_dispatch<5>(p, it, e, half); break;
        case 6: epilogue_dispatch<6>(p, it, e, half); break;
        default: epilogue_dispatch<3>(p, it, e, half); break;
      }
      if (p.out_mode == 1) {
        fence_proxy_async_smem();  // staging writes (generic proxy) -> bulk copy (async proxy)
        mbar_arrive(smem_u32(&ctl.out_full[2 * pair]));
      }
    }
    }
  } else if (warp == kStoreWarp) {
    // ===================================== output stores =====================================
    const int pair = tid & 31;
    if constexpr (VEC == kVecRaw9) {
      if (pair == 2) {
        // Raw-row producer of the 3x3x3 stem loader: requests the input rows of every item into the staging ring, as far
        // ahead as the ring allows.  (Round 2: this used to be loader thread 0, in front of its own share of every item;
        // decoding an item's row segments is ~400 serial instructions, and since a stage is complete only when all 128
        // loader threads have arrived, that lane's detour was ~2 us per item for the whole pipeline — measured as the
        // slope of the stem's time against its item count.)
        const uint32_t raw0 = smem_base + (uint32_t) p.smem_raw_off;
        int pb = 0;
        uint32_t ppar = 0;
        for (long long pitem = first; pitem < p.total_items; pitem += step) {
          const Item nx = decode_item(p, pitem);
          RawSeg sg[2];
          const int ns = raw_segments(p, nx, sg);
          mbar_wait_relaxed(smem_u32(&ctl.raw_empty[pb]), ppar ^ 1, 64);
          const uint32_t bar = smem_u32(&ctl.raw_full[pb]);
          mbar_arrive_expect_tx(bar, sg[0].bytes + (ns > 1 ? sg[1].bytes : 0u));
          for (int s = 0; s < ns; s++) {
            // several medium copies instead of one large one: they proceed in parallel
            for (uint32_t o = 0; o < sg[s].bytes; o += 4096) {
              const uint32_t len = sg[s].bytes - o < 4096 ? sg[s].bytes - o : 4096;
              bulk_g2s(raw0 + (uint32_t) pb * (uint32_t) p.raw_cap + sg[s].soff + o, sg[s].g + o, len, bar);
            }
          }
          if (++pb == p.raw_bufs) pb = 0, ppar ^= 1;
        }
      }
    }
    if (pair < 2 && p.out_mode == 2) {
      // panel epilogue: every item (tails and ragged n-tiles included) leaves through 2-D tensor stores, one per panel
      // and box of rows; the TMA unit undoes the panel swizzle and clips rows >= M / columns >= N
      const uint32_t staging0 = smem_base + p.smem_stage_off + pair * p.staging_bufs * p.staging_bytes;
      uint32_t k = 0;
      for (long long item = first + pair * step; item < p.total_items; item += 2 * step, k++) {
        const Item it = decode_item(p, item);
        const uint32_t buf = p.staging_bufs == 2 ? (k & 1u) : 0u, use = p.staging_bufs == 2 ? (k >> 1) : k;
        const uint32_t staging = staging0 + buf * (uint32_t) p.staging_bytes;
        mbar_wait_relaxed(smem_u32(&ctl.out_full[2 * pair + buf]), use & 1, 20);
        const int rows = it.mt_eff * kTileM;
        if (p.e2_dense) {  // contiguous output rows: the whole item is one run of bytes
          const long long left = p.M - it.m0;
          const uint32_t valid = left < rows ? (uint32_t) left : (uint32_t) rows;
          bulk_s2g(p.out + (size_t) it.m0 * p.out_stride, staging, valid * (uint32_t) p.goc);
        }
        for (int pk = 0; pk < p.e2_panels; pk++) {
          const int col = it.nt * p.n_tile + p.e2_col0[pk];
          if (col >= p.goc) break;
          const void* map = &smaps.m[p.e2_map[pk]][0];
          for (int r0 = 0; r0 < rows; r0 += p.e2_box_rows)
            tma_store_2d(map, staging + (uint32_t) p.e2_off[pk] + (uint32_t) (r0 * p.e2_width[pk]), col, (int) (it.m0 + r0));
        }
        bulk_commit();
        // the buffer is handed back as soon as its stores have read it.  With two buffers the pair does not need it before
        // the item after next, so this wait (~1 us) is off everybody's critical path.  (A first version released a buffer
        // one item LATE — after the next item's stores had been issued — and the final profile still showed the epilogue
        // warps waiting 5 % of their time for it.)
        bulk_wait_read<0>();
        mbar_arrive(smem_u32(&ctl.out_free[2 * pair + buf]));
      }
      bulk_wait<0>();
    } else if (pair < 2 && p.out_mode == 1) {
      const uint32_t staging = smem_base + p.smem_stage_off + pair * p.staging_bufs * p.staging_bytes;
      uint32_t k = 0;
      for (long long item = first + pair * step; item < p.total_items; item += 2 * step, k++) {
        const Item it = decode_item(p, item);
        mbar_wait_relaxed(smem_u32(&ctl.out_full[2 * pair]), k & 1, 20);
        if (it.m0 + (long long) it.mt_eff * kTileM <= p.M) {
          bulk_s2g(p.out + (size_t) it.m0 * p.out_stride, staging, (uint32_t) (it.mt_eff * kTileM * p.goc));
          bulk_commit();
          bulk_wait_read<0>();
        }
        mbar_arrive(smem_u32(&ctl.out_free[2 * pair]));
      }
      bulk_wait<0>();
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// Raises the kernel's dynamic shared-memory limit to everything the device allows beside its static shared memory.
static cudaError_t set_max_dynamic_smem(const void* kern, int max_smem_optin) {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, kern);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin - (int) fa.sharedSizeBytes);
}

// ------------------------------------------------------------------------------------------------
// large q8gemm on CTA pairs (cta_group::2)
// ------------------------------------------------------------------------------------------------
// For GEMMs whose weights do not fit shared memory (the tensor-bound regime: BASELINE.json's "int8 TOPS & %-of-peak on
// q8gemm") the 128 x 256 tiles of the kernel above pull (128 + 256) K bytes from L2 per 128 x 256 x K MACs — at the
// ~32 B/clk/SM the L2 delivered in that run this caps the tensor pipe near a third of its peak (measured 1.47 POPS of
// 4.55).  Here two CTAs of a cluster share every UMMA: M = 256 (128 rows per CTA), N = 256 with each CTA holding HALF of
// the B tile, so a CTA moves (128 + 128) K bytes for the same MACs — 1.5x less — and issues half the instructions.
//   * operands: TMA 2-D boxes of 128 rows x 128 bytes of K with the 128-byte swizzle (the K-major layout UMMA reads
//     without bank conflicts); A straight from the caller's activation matrix, B from a packed copy of the weights
//     [n-tile][256 rows][K]: 240 output channels, one all-ones row (row sums of A: the reference's own XZP algebra,
//     src/q8gemm/4x8c2-xzp-neon.c:26-67, pack.h:216-232) and 15 zero rows per tile;
//   * both CTAs' loads report to the LEADER's `full` barrier (it alone issues the UMMAs); tcgen05.commit multicasts
//     `empty` / `tmem_full` to both CTAs; epilogue warps of both CTAs release an accumulator stage on the leader's barrier;
//   * epilogue = the panel epilogue above ("ones" form), per CTA for its own 128 rows, double-buffered in TMEM (2 x 256
//     columns) and in the staging panels, so it overlaps the next tile's UMMAs completely.
constexpr int k2Stages = 5;                 // 5 x (16 KB A + 16 KB B half) = 160 KB ring
constexpr int k2StageBytes = 32768;
constexpr int k2NTile = 240;                // output channels per tile (+ 16 rows: ones / zero) = UMMA N 256
constexpr int k2EpiWarp0 = 2, k2EpiWarps = 8;
constexpr int k2StoreWarp = k2EpiWarp0 + k2EpiWarps;
constexpr int k2Threads = (k2StoreWarp + 1) * 32;  // 352

struct __align__(8) Gemm2Ctl {
  uint64_t full[k2Stages];
  uint64_t empty[k2Stages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];   // used in the leader CTA only: 2 CTAs x 8 epilogue warps arrive
  uint64_t out_full[2];
  uint64_t out_free[2];
  uint32_t tmem_base;
};

template <int RQ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads, 1)
    q8_gemm2sm_kernel(const __grid_constant__ IgemmParams p, const __grid_constant__ CUtensorMap tmap_a,
                      const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ IgemmStoreMaps smaps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ Gemm2Ctl ctl;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ring = smem_base;                                     // k2Stages x 32 KB
  const uint32_t staging0 = ring + k2Stages * k2StageBytes;            // 2 x staging_bytes (1024-aligned)
  const uint32_t bias_slot = staging0 + 2 * (uint32_t) p.staging_bytes;  // 2 x 1 KB: the tile's folded biases

  if (tid == 0) {
    for (int s = 0; s < k2Stages; s++) {
      mbar_init(smem_u32(&ctl.full[s]), 1);
      mbar_init(smem_u32(&ctl.empty[s]), 1);
    }
    for (int s = 0; s < 2; s++) {
      mbar_init(smem_u32(&ctl.tmem_full[s]), 1);
      mbar_init(smem_u32(&ctl.tmem_empty[s]), 2 * k2EpiWarps);
      mbar_init(smem_u32(&ctl.out_full[s]), k2EpiWarps * 32);
      mbar_init(smem_u32(&ctl.out_free[s]), 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(smem_u32(&ctl.tmem_base));
  tc_fence_before_sync();
  cluster_sync_all();   // barriers of BOTH CTAs are initialised before anyone signals across the pair
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl.tmem_base;

  const int n_tiles = p.n_tiles;
  const long long m_pairs = (p.M + 255) / 256;
  const long long tiles = m_pairs * n_tiles;
  const long long first = blockIdx.x >> 1, step = gridDim.x >> 1;
  const int num_kb = (p.K + 127) / 128;

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) =====================================
    if (lane == 0) {
      const uint32_t lead_full0 = mapa_shared(smem_u32(&ctl.full[0]), 0);
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = first; tile < tiles; tile += step) {
        const long long mp = tile / n_tiles;
        const int nt = (int) (tile - mp * n_tiles);
        const int row_a = (int) (mp * 256 + rank * 128);
        // the pair's B operand has n_mma rows, CTA r supplying rows [r * n_mma / 2, (r + 1) * n_mma / 2): the last (ragged)
        // n-tile runs a narrower UMMA (its real columns rounded to 16, plus the 16 rows holding the ones row)
        const int n_mma = nt == n_tiles - 1 ? p.e2_last_nmma : 256;
        const int row_b = nt * 256 + (int) rank * (n_mma >> 1);
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait_relaxed(smem_u32(&ctl.empty[stage]), phase ^ 1, 32);
          const uint32_t dst = ring + (uint32_t) stage * k2StageBytes;
          const uint32_t bar = lead_full0 + 8u * (uint32_t) stage;
          if (rank == 0) mbar_arrive_expect_tx(smem_u32(&ctl.full[stage]), 2u * k2StageBytes);  // both CTAs' bytes
          tma_load_2d_2sm(dst, &tmap_a, kb * 128, row_a, bar);
          tma_load_2d_2sm(dst + 16384, &tmap_b, kb * 128, row_b, bar);
          if (++stage == k2Stages) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== UMMA issue (leader CTA) =====================================
    if (rank == 0) {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t ring_u = __shfl_sync(0xffffffffu, ring, 0);
      const uint32_t ctl_u = __shfl_sync(0xffffffffu, smem_u32(&ctl), 0);
      const uint32_t bar_full = ctl_u + (uint32_t) offsetof(Gemm2Ctl, full), bar_empty = ctl_u + (uint32_t) offsetof(Gemm2Ctl, empty);
      const uint32_t bar_tfull = ctl_u + (uint32_t) offsetof(Gemm2Ctl, tmem_full);
      const uint32_t bar_tempty = ctl_u + (uint32_t) offsetof(Gemm2Ctl, tmem_empty);
      const uint32_t idesc_full = umma_idesc_i8(256, 256, false, false);  // u8 x u8 ("ones" algebra), M = 256 over the pair
      const uint32_t idesc_last = umma_idesc_i8(256, (uint32_t) p.e2_last_nmma, false, false);
      int stage = 0, as = 0;
      uint32_t phase = 0, as_phase = 0;
      for (long long tile = first; tile < tiles; tile += step) {
        const uint32_t idesc = (int) (tile % n_tiles) == n_tiles - 1 ? idesc_last : idesc_full;
        mbar_wait_parked(bar_tempty + 8u * (uint32_t) as, as_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_u + (uint32_t) as * 256u;
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait_parked(bar_full + 8u * (uint32_t) stage, phase);
          tc_fence_after_sync();
          const uint32_t a_addr = ring_u + (uint32_t) stage * k2StageBytes;
          if (elect_one()) {
            const uint64_t ad = umma_desc_kmajor_sw128(a_addr), bd = umma_desc_kmajor_sw128(a_addr + 16384);
#pragma unroll
            for (int k4 = 0; k4 < 4; k4++)  // 32 bytes of K per UMMA: +2 in the 16-byte-granular start address
              umma_i8_2sm(d_tmem, ad + (uint64_t) (2 * k4), bd + (uint64_t) (2 * k4), idesc, (kb | k4) != 0 ? 1u : 0u);
            umma_commit_2sm(bar_empty + 8u * (uint32_t) stage, 3);
            if (kb == num_kb - 1) umma_commit_2sm(bar_tfull + 8u * (uint32_t) as, 3);
          }
          __syncwarp();
          if (++stage == k2Stages) stage = 0, phase ^= 1;
        }
        as ^= 1;
        if (as == 0) as_phase ^= 1;
      }
    }
  } else if (warp < k2StoreWarp) {
    // ===================================== epilogue (8 warps per CTA) =====================================
    const int ew = warp - k2EpiWarp0;
    const int q = warp & 3, half = ew >> 2;   // TMEM lane quarter = warp index mod 4 (hardware rule)
    const uint32_t row = (uint32_t) (q * 32 + lane);
    const uint32_t lead_tempty0 = mapa_shared(smem_u32(&ctl.tmem_empty[0]), 0);
    int as = 0;
    uint32_t as_phase = 0, k = 0;
    for (long long tile = first; tile < tiles; tile += step, k++) {
      const long long mp = tile / n_tiles;
      const int nt = (int) (tile - mp * n_tiles);
      const uint32_t buf = k & 1;
      // this tile's folded biases -> smem slot (the previous user of the slot, tile k - 2, is long done: its epilogue
      // finished before tile k - 1's, and every warp passes the named barrier of tile k - 1 before reaching this one)
      const uint32_t slot = bias_slot + buf * 1024u;
      for (int i = ew * 32 + lane; i < k2NTile; i += k2EpiWarps * 32)
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(slot + 4u * (uint32_t) i), "r"(__ldg(p.bias + (size_t) nt * k2NTile + i)) : "memory");
      named_bar_sync(1, k2EpiWarps * 32);
      mbar_wait(smem_u32(&ctl.tmem_full[as]), as_phase);
      tc_fence_after_sync();
      epi2_item<RQ, false, true>(p, 1, tmem_base + (uint32_t) as * 256u + ((uint32_t) (q * 32) << 16), slot,
                                 staging0 + buf * (uint32_t) p.staging_bytes, row, half, lead_tempty0 + 8u * (uint32_t) as,
                                 smem_u32(&ctl.out_free[buf]), ((k >> 1) & 1) ^ 1, nt == n_tiles - 1 ? p.e2_last_nmma - 16 : k2NTile);
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&ctl.out_full[buf]));
      as ^= 1;
      if (as == 0) as_phase ^= 1;
    }
  } else {
    // ===================================== output stores =====================================
    if (lane == 0) {
      uint32_t k = 0;
      for (long long tile = first; tile < tiles; tile += step, k++) {
        const long long mp = tile / n_tiles;
        const int nt = (int) (tile - mp * n_tiles);
        const uint32_t buf = k & 1;
        mbar_wait_relaxed(smem_u32(&ctl.out_full[buf]), (k >> 1) & 1, 20);
        const uint32_t staging = staging0 + buf * (uint32_t) p.staging_bytes;
        const long long m0 = mp * 256 + rank * 128;
        for (int pk = 0; pk < p.e2_panels; pk++) {
          const int col = nt * k2NTile + p.e2_col0[pk];
          if (col >= p.goc || m0 >= p.M) break;
          tma_store_2d(&smaps.m[p.e2_map[pk]][0], staging + (uint32_t) p.e2_off[pk], col, (int) m0);
        }
        bulk_commit();
        bulk_wait_read<0>();
        mbar_arrive(smem_u32(&ctl.out_free[buf]));
      }
      bulk_wait<0>();
    }
  }

  tc_fence_before_sync();
  cluster_sync_all();   // both CTAs are done with the pair's tensor memory and with each other's barriers
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

cudaError_t launch_q8_gemm2sm(const IgemmParams& p, const void* tmap_a, const void* tmap_b, const IgemmStoreMaps& smaps, int clusters,
                              int max_smem_optin, cudaStream_t stream) {
  alignas(64) CUtensorMap ta, tb;
  memcpy(&ta, tmap_a, sizeof(ta));
  memcpy(&tb, tmap_b, sizeof(tb));
  const int smem = k2Stages * k2StageBytes + 2 * p.staging_bytes + 2048 + 1024;
  if (p.rq_mode == 5) {
    auto kern = q8_gemm2sm_kernel<5>;
    static cudaError_t attr_status = set_max_dynamic_smem(reinterpret_cast<const void*>(kern), max_smem_optin);
    if (attr_status != cudaSuccess) return attr_status;
    kern<<<2 * clusters, k2Threads, smem, stream>>>(p, ta, tb, smaps);
  } else {
    auto kern = q8_gemm2sm_kernel<6>;
    static cudaError_t attr_status = set_max_dynamic_smem(reinterpret_cast<const void*>(kern), max_smem_optin);
    if (attr_status != cudaSuccess) return attr_status;
    kern<<<2 * clusters, k2Threads, smem, stream>>>(p, ta, tb, smaps);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
// The dynamic-smem limit is a property of the kernel instantiation, not of a launch: it is raised ONCE to the device's
// opt-in maximum (setting it per launch to the operator's size raced between host threads running different operators).
template <int MODE, int VEC>
static cudaError_t launch_one(const IgemmParams& p, const CUtensorMap& tmap, const IgemmStoreMaps& smaps, int grid,
                              int max_smem_optin, cudaStream_t stream) {
  auto kern = q8_igemm_kernel<MODE, VEC>;
  static cudaError_t attr_status = set_max_dynamic_smem(reinterpret_cast<const void*>(kern), max_smem_optin);
  if (attr_status != cudaSuccess) return attr_status;
  kern<<<grid, kThreads, p.smem_total, stream>>>(p, tmap, smaps);
  return cudaGetLastError();
}

// vec: 16/8/4/1 = cp.async piece size, 0 = 9-byte row runs (3x3 over 3 channels), 32 = TMA (gemm mode; `tmap` valid)
cudaError_t launch_q8_igemm(const IgemmParams& p, int mode, int vec, const void* tmap_a, const IgemmStoreMaps* store_maps,
                            int grid, int max_smem_optin, cudaStream_t stream) {
  alignas(64) CUtensorMap tmap;
  if (tmap_a != nullptr) {
    memcpy(&tmap, tmap_a, sizeof(tmap));
  } else {
    memset(&tmap, 0, sizeof(tmap));
  }
  static const IgemmStoreMaps no_maps{};
  const IgemmStoreMaps& sm = store_maps != nullptr ? *store_maps : no_maps;
#define Q8_LAUNCH(M, V) return launch_one<M, V>(p, tmap, sm, grid, max_smem_optin, stream)
  if (mode == kModeGemm) {
    switch (vec) {
      case 32: Q8_LAUNCH(kModeGemm, 32);
      case 16: Q8_LAUNCH(kModeGemm, 16);
      case 8: Q8_LAUNCH(kModeGemm, 8);
      case 4: Q8_LAUNCH(kModeGemm, 4);
      default: Q8_LAUNCH(kModeGemm, 1);
    }
  } else {
    switch (vec) {
      case 16: Q8_LAUNCH(kModeConv, 16);
      case 8: Q8_LAUNCH(kModeConv, 8);
      case 4: Q8_LAUNCH(kModeConv, 4);
      case 0: Q8_LAUNCH(kModeConv, 0);         // 9-byte row runs (3x3, 3 channels)
      case 2: Q8_LAUNCH(kModeConv, kVecRaw9);  // ... from bulk-staged raw rows
      default: Q8_LAUNCH(kModeConv, 1);
    }
  }
#undef Q8_LAUNCH
}

}  // namespace q8
