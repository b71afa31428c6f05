// q8gemm / q8conv for sm_100a: persistent, warp-specialised implicit GEMM on the 5th-gen tensor cores.
//
// Replaces (reference, paths relative to its root):
//   src/operator-run.c:770-804 + :39-70   gemm case  -> q8gemm_ukernel_4x4c2__sse2 (src/q8gemm/4x4c2-sse2.c:14-318)
//   src/operator-run.c:805-844 + :183-217 conv case  -> q8conv_ukernel_4x4c2__sse2 (src/q8conv/4x4c2-sse2.c:14-273)
//   src/indirection.c:18-79 (the pointer table is never built: tap -> address is computed in the load stage)
//   the Q31 epilogue those micro-kernels inline (src/q8gemm/4x4c2-sse2.c:178-278)
//
// Arithmetic (bit-exact by construction, all int32, order independent):
//   acc[m][n] = bias'[n] + sum_k a[m][k] * w[n][k] - kzp * sum_k a[m][k]
// with bias' = b + K*izp*kzp - izp*sum_k w (the reference's packed bias) and padded taps reading the
// byte izp, i.e. the reference's own "XZP" algebra (src/q8gemm/4x8c2-xzp-neon.c:26-67, pack.h:216-232)
// which lets the tensor core run raw u8 x u8 -> s32.  sum_k a[m][k] falls out of the same UMMA as one
// extra B row of ones (accumulator column n_tile).
//
// Structure (one CTA per SM, 672 threads):
//   warps 0-7   epilogue pair 0 (TMEM stage 0, even work items)   TMEM -> regs -> Q31 requant -> uint8 -> global
//   warps 8-15  epilogue pair 1 (TMEM stage 1, odd work items)
//   warp  16    TMEM allocation + UMMA issue (one lane)
//   warps 17-20 loaders: cp.async global -> smem, canonical K-major no-swizzle layout [sub-tile][k-chunk][row][16 B]
// A work item is `mt` (<= 8) consecutive 128-row sub-tiles x one n-tile: the sub-tiles' accumulators sit side
// by side in one TMEM stage (mt * n_mma <= 256 columns), which amortises every per-item synchronisation
// over up to 1024 rows — essential for the narrow (N = 16..96) projection layers.
// Pipelines: smem ring full/empty mbarriers (loaders <-> UMMA), two TMEM accumulator stages
// full/empty (UMMA <-> epilogue pairs).  int32 accumulators never leave TMEM/registers.
#include <cuda_runtime.h>
#include <stdint.h>

#include "q8_igemm_sm100.cuh"
#include "sm100_ptx.cuh"

namespace q8 {

constexpr int kEpiWarps = 16;
constexpr int kEpiPairThreads = 256;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kLoadWarp0 = kMmaWarp + 1;
constexpr int kLoadWarps = 4;
constexpr int kLoadThreads = kLoadWarps * 32;
constexpr int kThreads = (kLoadWarp0 + kLoadWarps) * 32;  // 672
constexpr int kTmemCols = 512;

struct __align__(8) SmemCtl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t b_full;
  uint32_t tmem_base;
};

struct Item {
  long long m0;  // first row of the item
  int g, nt;
  int mt_eff;    // sub-tiles that contain at least one valid row
};

__device__ __forceinline__ Item decode_item(const IgemmParams& p, long long item) {
  Item it;
  it.nt = (int) (item % p.n_tiles);
  const long long rest = item / p.n_tiles;
  const long long st = rest % p.m_super;
  it.g = (int) (rest / p.m_super);
  it.m0 = st * p.mt * kTileM;
  const long long left = p.m_tiles - st * p.mt;
  it.mt_eff = left < p.mt ? (int) left : p.mt;
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
      const int ox = (int) (m % p.out_w);
      const long long t = m / p.out_w;
      const int oy = (int) (t % p.out_h);
      const long long n = t / p.out_h;
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

__device__ __forceinline__ void copy_bytes16(uint32_t dst, const uint8_t* src, int bytes, int ltid) {
  for (int o = ltid * 16; o < bytes; o += kLoadThreads * 16) cp_async<16>(dst + o, src + o);
}

// ------------------------------------------------------------------------------------------------
// epilogue
// ------------------------------------------------------------------------------------------------
template <int RQ>
__device__ __forceinline__ uint32_t requant4(const int32_t* v, const IgemmParams& p, const int4 b, int32_t corr) {
  int32_t y[4];
  const int32_t bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int32_t n = v[i] + bb[i] + corr;
    if constexpr (RQ == 0) {
      y[i] = q8_requant_fused_unclamped(n, p.rq.multiplier, p.rq.c_pos, p.rq.shift - 1);
    } else if constexpr (RQ == 1) {
      int32_t t = q8_requant_fused_unclamped(n, p.rq.multiplier, p.rq.c_pos, p.rq.shift - 1);
      t = max(t, p.rq.qmin);
      y[i] = min(t, p.rq.qmax);
    } else if constexpr (RQ == 2) {
      y[i] = q8_requant_shift0(n, p.rq.multiplier, p.rq.zero_point, p.rq.qmin, p.rq.qmax);
    } else if constexpr (RQ == 4) {
      int32_t t = q8_requant_fused_shift1_unclamped(n, p.rq.multiplier, p.rq.c_neg);
      t = max(t, p.rq.qmin);
      y[i] = min(t, p.rq.qmax);
    } else {
      y[i] = q8_requant_exact_slow(n, p.rq);
    }
  }
  return pack_sat_u8x4(y[0], y[1], y[2], y[3]);  // saturation to [0,255] is the clamp when qmin=0,qmax=255
}

// Slow path of the direct store: fewer than 16 valid bytes, or a destination that is not 16-byte aligned.
__device__ __noinline__ void store_row_partial(uint8_t* dst, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int valid,
                                               int vec) {
  const uint32_t w[4] = {w0, w1, w2, w3};
  int i = 0;
  if (vec >= 4) {
    for (; i + 4 <= valid; i += 4) *reinterpret_cast<uint32_t*>(dst + i) = w[i >> 2];
  }
  for (; i < valid; i++) dst[i] = (uint8_t) (w[i >> 2] >> (8 * (i & 3)));
}

__device__ __noinline__ void dump_acc(const IgemmParams& p, long long item, int j, int row, int c0, const int32_t* v,
                                      int32_t rowsum) {
  int32_t* d = p.dbg_acc + (((size_t) item * p.mt + j) * kTileM + row) * p.n_mma;
  for (int i = 0; i < 16; i++) d[c0 + i] = v[i];
  if (c0 == 0 && p.has_corr) d[p.n_tile] = rowsum;
}

// One epilogue warp: lane quarter q = warp % 4 of the accumulator, every second (sub-tile, 16-column) unit.
template <int RQ>
__device__ __forceinline__ void epilogue_item(
    const IgemmParams& p, const Item& it, long long item, uint32_t tmem_acc, int q, int half, int lane, uint32_t bias_smem,
    uint32_t staging, bool bulk) {
  const int row = q * 32 + lane;  // TMEM lane == row inside a sub-tile
  const uint32_t tlane = tmem_acc + ((uint32_t) (q * 32) << 16);
  const int ch = p.n_tile >> 4;   // 16-column chunks per sub-tile
  const int units = it.mt_eff * ch;
  const int n_valid = min(p.n_tile, p.goc - it.nt * p.n_tile);
  const uint32_t bias_base = bias_smem + (uint32_t) ((it.g * p.n_tiles + it.nt) * p.n_tile) * 4;
  uint8_t* const obase = p.out + (size_t) it.g * p.goc + (size_t) it.nt * p.n_tile;

  int j = half / ch, c = half - j * ch;
  for (int u = half; u < units; u += 2) {
    const int c0 = c << 4;
    int32_t v[16];
    int32_t rowsum = 0;
    tmem_ld16(tlane + j * p.n_mma + c0, v);
    if (p.has_corr) tmem_ld1(tlane + j * p.n_mma + p.n_tile, rowsum);
    int4 b[4];
#pragma unroll
    for (int t = 0; t < 4; t++)
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(b[t].x), "=r"(b[t].y), "=r"(b[t].z), "=r"(b[t].w)
                   : "r"(bias_base + (uint32_t) (c0 + 4 * t) * 4));
    tmem_ld_wait();
    if (p.dbg_acc != nullptr) dump_acc(p, item, j, row, c0, v, rowsum);
    const int32_t corr = -p.kzp * rowsum;
    uint32_t w[4];
#pragma unroll
    for (int t = 0; t < 4; t++) w[t] = requant4<RQ>(v + 4 * t, p, b[t], corr);

    const int valid = n_valid - c0;
    if (valid > 0) {
      if (bulk) {
        // staging = dense image of the item's output rows (pitch goc; goc % 4 == 0 guaranteed by the host)
        const uint32_t s = staging + (uint32_t) (j * kTileM + row) * p.goc + c0;
        if (valid >= 16 && (p.goc & 15) == 0) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                       : "memory");
        } else {
#pragma unroll
          for (int t = 0; t < 4; t++)
            if (4 * t < valid) asm volatile("st.shared.b32 [%0], %1;" ::"r"(s + 4 * t), "r"(w[t]) : "memory");
        }
      } else {
        const long long m = it.m0 + (long long) j * kTileM + row;
        if (m < p.M) {
          uint8_t* dst = obase + (size_t) m * p.out_stride + c0;
          if (valid >= 16 && p.out_vec == 16) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
          } else {
            store_row_partial(dst, w[0], w[1], w[2], w[3], valid < 16 ? valid : 16, p.out_vec);
          }
        }
      }
    }
    c += 2;
    while (c >= ch) {
      c -= ch;
      ++j;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int MODE, int VEC>
__global__ void __launch_bounds__(kThreads, 1) q8_igemm_kernel(const __grid_constant__ IgemmParams p) {
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
      mbar_init(smem_u32(&ctl.full[s]), kLoadThreads);
      mbar_init(smem_u32(&ctl.empty[s]), 1);
    }
    for (int s = 0; s < 2; s++) {
      mbar_init(smem_u32(&ctl.tmem_full[s]), 1);
      mbar_init(smem_u32(&ctl.tmem_empty[s]), kEpiPairThreads);
    }
    mbar_init(smem_u32(&ctl.b_full), kLoadThreads);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc<kTmemCols>(smem_u32(&ctl.tmem_base));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl.tmem_base;

  const long long first = blockIdx.x, step = gridDim.x;

  if (warp >= kLoadWarp0) {
    // ===================================== loaders =====================================
    const int ltid = tid - kLoadWarp0 * 32;
    {
      // one-time: folded biases (always) and the packed weights (when they fit) become smem-resident
      copy_bytes16(bias_smem, reinterpret_cast<const uint8_t*>(p.bias), p.bias_count * 4, ltid);
      if (p.b_resident) copy_bytes16(b_smem, p.wpack, p.groups * p.n_tiles * p.nkc * p.n_mma * 16, ltid);
      cp_async_mbar_arrive_noinc(smem_u32(&ctl.b_full));
    }
    int stage = 0;
    uint32_t phase = 0;
    for (long long item = first; item < p.total_items; item += step) {
      const Item it = decode_item(p, item);
      for (int ks = 0; ks < p.k_stages; ks++) {
        mbar_wait(smem_u32(&ctl.empty[stage]), phase ^ 1);
        const uint32_t a_stage = a_smem + stage * p.stage_bytes;
        if constexpr (MODE == kModeGemm) {
          load_a_gemm<VEC>(p, it, ks, a_stage, ltid);
        } else {
          load_a_conv<VEC>(p, it, ks, a_stage, ltid);
        }
        if (!p.b_resident) {
          int cs = p.nkc - ks * p.skc;
          cs = cs < p.skc ? cs : p.skc;
          const uint8_t* wsrc =
              p.wpack + ((size_t) (it.g * p.n_tiles + it.nt) * p.nkc + (size_t) ks * p.skc) * p.n_mma * 16;
          copy_bytes16(a_stage + p.mt * p.skc * kChunkBytes, wsrc, cs * p.n_mma * 16, ltid);
        }
        fence_proxy_async_smem();  // st.shared fills (padding taps / byte path) -> UMMA reads
        cp_async_mbar_arrive_noinc(smem_u32(&ctl.full[stage]));
        if (++stage == p.num_stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    cp_async_wait_all();
  } else if (warp == kMmaWarp) {
    // ===================================== UMMA issue =====================================
    if ((tid & 31) == 0) {
      const uint32_t idesc = umma_idesc_i8(kTileM, (uint32_t) p.n_mma, false, false);
      const uint32_t b_lbo = (uint32_t) p.n_mma * 16;
      const uint32_t sub_bytes = (uint32_t) p.skc * kChunkBytes;
      if (p.b_resident) {
        mbar_wait(smem_u32(&ctl.b_full), 0);
        fence_proxy_async_smem();
      }
      int stage = 0;
      uint32_t phase = 0;
      long long li = 0;
      for (long long item = first; item < p.total_items; item += step, li++) {
        const Item it = decode_item(p, item);
        const int as = (int) (li & 1);
        mbar_wait(smem_u32(&ctl.tmem_empty[as]), (uint32_t) ((li >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * kMaxNMma;
        for (int ks = 0; ks < p.k_stages; ks++) {
          mbar_wait(smem_u32(&ctl.full[stage]), phase);
          fence_proxy_async_smem();
          tc_fence_after_sync();
          const uint32_t a_stage = a_smem + stage * p.stage_bytes;
          int cs = p.nkc - ks * p.skc;
          cs = cs < p.skc ? cs : p.skc;
          const uint32_t b_base = p.b_resident
              ? b_smem + (uint32_t) (((it.g * p.n_tiles + it.nt) * p.nkc + ks * p.skc) * p.n_mma * 16)
              : a_stage + p.mt * sub_bytes;
          for (int j = 0; j < it.mt_eff; j++) {
            for (int c = 0; c < cs; c += 2) {
              const uint64_t a_desc = umma_desc_kmajor_noswizzle(a_stage + j * sub_bytes + c * kChunkBytes, kChunkBytes, 128);
              const uint64_t b_desc = umma_desc_kmajor_noswizzle(b_base + c * b_lbo, b_lbo, 128);
              umma_i8(d_tmem + j * p.n_mma, a_desc, b_desc, idesc, (ks | c) != 0 ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&ctl.empty[stage]));
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(smem_u32(&ctl.tmem_full[as]));
      }
    }
  } else {
    // ===================================== epilogue =====================================
    const int pair = warp >> 3;          // 0 or 1 == TMEM stage
    const int pw = warp & 7;
    const int q = pw & 3, half = pw >> 2;
    const int lane = tid & 31;
    const int pair_tid = tid & (kEpiPairThreads - 1);
    const uint32_t staging = smem_base + p.smem_stage_off + pair * p.staging_bytes;
    mbar_wait(smem_u32(&ctl.b_full), 0);  // biases are in smem
    long long li = pair;
    bool bulk_pending = false;
    for (long long item = first + pair * step; item < p.total_items; item += 2 * step, li += 2) {
      const Item it = decode_item(p, item);
      const bool bulk = p.out_mode == 1 && (it.m0 + (long long) it.mt_eff * kTileM <= p.M);
      if (p.out_mode == 1) {
        // the previous bulk store of this pair must have finished reading the staging buffer
        if (pair_tid == 0 && bulk_pending) bulk_wait_read<0>();
        named_bar_sync(1 + pair, kEpiPairThreads);
      }
      mbar_wait(smem_u32(&ctl.tmem_full[pair]), (uint32_t) ((li >> 1) & 1));
      tc_fence_after_sync();
      const uint32_t tmem_acc = tmem_base + pair * kMaxNMma;
      switch (p.rq_mode) {
        case 0: epilogue_item<0>(p, it, item, tmem_acc, q, half, lane, bias_smem, staging, bulk); break;
        case 1: epilogue_item<1>(p, it, item, tmem_acc, q, half, lane, bias_smem, staging, bulk); break;
        case 2: epilogue_item<2>(p, it, item, tmem_acc, q, half, lane, bias_smem, staging, bulk); break;
        case 4: epilogue_item<4>(p, it, item, tmem_acc, q, half, lane, bias_smem, staging, bulk); break;
        default: epilogue_item<3>(p, it, item, tmem_acc, q, half, lane, bias_smem, staging, bulk); break;
      }
      // the accumulator stage may be overwritten by the next-but-one work item
      tc_fence_before_sync();
      mbar_arrive(smem_u32(&ctl.tmem_empty[pair]));
      if (p.out_mode == 1) {
        fence_proxy_async_smem();
        named_bar_sync(1 + pair, kEpiPairThreads);
        if (bulk && pair_tid == 0) {
          bulk_s2g(p.out + (size_t) it.m0 * p.out_stride, staging, (uint32_t) (it.mt_eff * kTileM * p.goc));
          bulk_commit();
          bulk_pending = true;
        }
      }
    }
    if (pair_tid == 0 && bulk_pending) bulk_wait<0>();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
template <int MODE, int VEC>
static cudaError_t launch_one(const IgemmParams& p, int grid, cudaStream_t stream) {
  auto kern = q8_igemm_kernel<MODE, VEC>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_total);
  if (e != cudaSuccess) return e;
  kern<<<grid, kThreads, p.smem_total, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_q8_igemm(const IgemmParams& p, int mode, int vec, int grid, cudaStream_t stream) {
  if (mode == kModeGemm) {
    switch (vec) {
      case 16: return launch_one<kModeGemm, 16>(p, grid, stream);
      case 8: return launch_one<kModeGemm, 8>(p, grid, stream);
      case 4: return launch_one<kModeGemm, 4>(p, grid, stream);
      default: return launch_one<kModeGemm, 1>(p, grid, stream);
    }
  } else {
    switch (vec) {
      case 16: return launch_one<kModeConv, 16>(p, grid, stream);
      case 8: return launch_one<kModeConv, 8>(p, grid, stream);
      case 4: return launch_one<kModeConv, 4>(p, grid, stream);
      default: return launch_one<kModeConv, 1>(p, grid, stream);
    }
  }
}

}  // namespace q8
