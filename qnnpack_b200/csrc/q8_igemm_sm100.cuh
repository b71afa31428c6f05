// Parameters shared by the host launcher and the tcgen05 implicit-GEMM kernel (q8gemm + q8conv).
#pragma once
#include <stdint.h>

#include "requant_math.h"

namespace q8 {

constexpr int kTileM = 128;               // rows (output pixels) per UMMA = TMEM lanes
constexpr int kChunkBytes = kTileM * 16;  // one 16-byte K-chunk of an A sub-tile: [128 rows][16 B]
constexpr int kMaxStages = 16;
constexpr int kMaxNMma = 256;             // UMMA N limit; also the TMEM column budget of one accumulator stage
constexpr int kOnesCols = 16;             // extra B rows: row 0 of the block is all-ones -> per-row sum of A
constexpr int kMaxNTile = kMaxNMma - kOnesCols;
constexpr int kMaxSubTiles = 8;
constexpr int kMaxAccStages = 4;          // TMEM accumulator stages (512 columns / (mt * n_mma), at most 4)           // 128-row sub-tiles per work item (they share one TMEM stage)

enum IgemmMode : int { kModeGemm = 0, kModeConv = 1 };

struct IgemmParams {
  const uint8_t* in;
  uint8_t* out;
  const uint8_t* wpack;  // [group][n_tile] blocks, each [nkc][n_mma][16 B]  (K-major, no-swizzle core matrices)
  const int32_t* bias;   // [group][n_tiles * n_tile] folded bias (reference pack.h:24,43,63,84)
  int32_t* dbg_acc;      // optional: raw accumulators [item][sub-tile][128][n_mma]

  long long M;           // rows per group = batch * out_h * out_w
  long long m_tiles;     // ceil(M / 128)
  long long m_super;     // ceil(m_tiles / mt)
  long long total_items; // groups * m_super * n_tiles
  long long in_stride, out_stride;
  int groups, gic, goc;

  // conv geometry (kModeConv)
  int in_h, in_w, out_h, out_w, kh, kw, stride_h, stride_w, dil_h, dil_w, pad_top, pad_left;

  // tiling
  int K;          // kh*kw*gic
  int nkc;        // 16-byte K chunks incl. padding, even
  int skc;        // chunks per pipeline stage, even
  int k_stages;   // ceil(nkc / skc)
  int mt;         // 128-row sub-tiles per work item
  int acc_stages; // TMEM accumulator stages in use (2..4)
  int acc_stride; // TMEM columns per accumulator stage (mt * n_mma)
  int n_tiles, n_tile, n_mma;
  int has_corr;   // "ones" mode only: B carries the ones block and the epilogue applies  - kzp * rowsum
  // "folded" mode: bias and zero-point correction are accumulated by extra UMMAs, the epilogue only requantises
  int folded;      // 1: folded mode (requires resident weights)
  int b_signed;    // folded: main B operand is (w - 128) as s8 (kzp != 0); 0: raw u8 weights (kzp == 0)
  int has_b2;      // folded: second UMMA per K step with the constant (128 - kzp) operand
  int bias_steps;  // folded: number of K=32 UMMA steps that add the folded bias (A = [255 x31, 1], B = s8 digits)
  int blk_chunks;  // 16-byte K chunks per (group, n_tile) block of `wpack`: nkc [+ 4 + 2*bias_steps when folded]
  int k_tail_pad;  // 1 if K < nkc*16 (last chunk pair partly padding -> uses the "tail" constant operand)
  int smem_aconst_off;
  int b_resident; // 1: all packed weights live in smem for the whole kernel
  int num_stages; // A(+B) ring depth
  int stage_bytes;
  int bias_count; // ints in `bias`
  int out_mode;   // 1: full items leave through the smem staging buffer + one bulk store; 0: per-thread global stores
  int out_vec;    // widest power-of-two (<= 32) dividing output base address, pixel stride and channel offsets
  int shift_mul;  // 2^(33 - shift) when the final shift can be done as a multiply-high (shift >= 3), else 0
  int rq_mode;    // 0: fused shift>=2, no clamp; 1: fused shift>=2 + clamp; 2: shift==0; 3: exact slow; 4: fused shift==1

  // smem carve-up (byte offsets into dynamic smem, 1024-aligned base)
  int smem_b_off, smem_bias_off, smem_a_off, smem_stage_off, staging_bytes, smem_total;
  int a_sw32;          // TMA loader: 1 = activations arrive as 32-byte K slabs in SWIZZLE_32B tiles (K % 32 == 0) instead of 16-byte
                       // chunks; 2 = K % 32 == 16 and one K stage: slabs for the first a_tail_c chunks, the last 16 bytes of K as a
                       // chunk pair (second chunk zero-filled) through the no-swizzle view in IgemmStoreMaps::m[4]
  int a_tail_c;
  int staging_bufs;    // output staging buffers per epilogue pair: 1, or 2 used alternately (panel epilogue, when smem allows)
  int smem_raw_off, raw_cap, raw_bufs;  // raw-row staging (3x3x3 stem loader): raw_bufs (2..4) buffers of raw_cap bytes
  int raw_batch;              // images in the input tensor (bounds the bulk copies)

  // out_mode 2 ("panel" epilogue): the pair's staging image is a set of column panels [panel][sub-tile*128 + row][width],
  // width in {128, 64, 32, 16} bytes with the matching TMA swizzle (conflict-free 16-byte staging stores for every
  // pitch), written to global memory by 2-D tensor stores that clip rows >= M and columns >= N themselves.
  int e2_dense;               // 1: one dense [row][N] image + ONE 1-D bulk store per item instead of panels — used when the
                              //    output rows are contiguous (stride == N) and the pitch N is conflict-free (N / 16 odd)
  int e2_panels;              // panels per n-tile (greedy split of n_tile into 128/64/32/16)
  int e2_box_rows;            // rows per tensor store: 256 when mt is even, else 128
  int e2_col0[4];             // first column of panel k inside the n-tile
  int e2_width[4];            // its width in bytes
  int e2_off[4];              // its byte offset inside the pair's staging buffer (1024-aligned)
  int e2_map[4];              // which tensor map (width class: 0 = 16, 1 = 32, 2 = 64, 3 = 128)
  // per epilogue unit c of a sub-tile (W = 32 columns folded / 16 otherwise):
  //   x = byte offset of the unit's first 16-byte chunk relative to the staging buffer, for row 0 of sub-tile 0
  //   y = pitch | lsh << 8 | mask << 16   (swizzle: chunk bits ^= (row << lsh) & mask)
  uint2 e2_unit[16];
  int e2_last_nmma;           // CTA-pair GEMM: UMMA N of the last n-tile (real columns rounded up to 16, + 16 for the ones row)

  int izp, kzp;
  Q8Requant rq;
};

// tensor maps of the output for the panel epilogue, one per panel-width class
struct IgemmStoreMaps {
  // [0..3]: output panels by width class; [4]: 16-byte-chunk view of the activations for the K tail (IgemmParams::a_sw32 == 2)
  alignas(64) unsigned char m[5][128];
};

}  // namespace q8
