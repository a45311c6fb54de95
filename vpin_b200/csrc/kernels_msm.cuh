// Fixed-base multi-scalar multiplication over ristretto255 for Hyrax/Pedersen commitments on sm_100a.
//
// Every MSM in the prover is over a prefix of one fixed generator stream (Spartan/src/commitments.rs:20-38), and a
// Hyrax commitment is L independent MSMs over the SAME R generators (Spartan/src/dense_mlpoly.rs:160-175). So instead
// of per-row Pippenger buckets (rows are only 2^8..2^15 points long) the device keeps, per generator G_j, kMsmSub tables:
// sub-table t holds the multiples 1..2^(W-1) of 2^(W*geom.group*t) G_j in affine Niels form with every coordinate halved,
// ((y+x)/2, (y-x)/2, d x y) - 96 B each; the mixed addition then uses D = Z1 instead of 2 Z1 (kernels_msm.cu). Scalars are
// recoded into signed W-bit digits; window w = t*geom.group + w' then contributes table_t[j][|digit|] to the partial sum
// of "local window" w'. One thread owns one (row, local window, column segment) and adds kMsmSub looked-up multiples per
// column with 7-multiplication mixed additions (no bucket reduction, no doublings in the hot loop; zero digits —
// addresses, timestamps, +-1 — are skipped). A finishing kernel adds the segment partials, runs a Horner pass over the
// geom.group local windows (5 x W doublings instead of 21 x W thanks to the sub-tables) and encodes the row.
// Results equal the reference's vartime_multiscalar_mul (Spartan/src/group.rs:103-121) as group elements, hence as
// compressed bytes.
#pragma once
#include <cuda_runtime.h>

#include "ed.cuh"

namespace vpin {

static const int kMsmSub = 4;                      // sub-tables per generator (the default; MsmGeom::sub is what a table was built with)
static const int kMsmColsPerBlock = 64;            // digit rows are padded to a multiple of this many columns
static const int kMsmRowsPerBlock = 128;           // threads per accumulate block (consecutive rows)
static const int kMsmMinW = 12, kMsmMaxW = 15;     // window widths the kernels support (digit = 15-bit magnitude | sign)
static const int kMsmMaxGroup = 10;                // longest Horner pass: two sub-tables at W = 13 (20 windows)

// Window geometry of one generator stream. The window width is chosen per stream when its table is built: the widest
// one whose table fits the memory budget (a wider window means fewer mixed additions per scalar — 22 at W = 12, 17 at
// W = 15 — for a table that doubles with every bit: 786 KB per generator at W = 12, 6.3 MB at W = 15).
// Sub-tables trade memory against the Horner pass: with `sub` tables per generator a row needs group = ceil(windows / sub) local
// windows and (group - 1) W doublings at the end. TWO sub-tables of width W + 1 take the memory of four of width W and save a
// window's worth of additions per scalar (W = 12 -> 13: 22 -> 20, 14 -> 15: 19 -> 17) for a Horner pass twice as long - the
// better deal whenever the HBM plan of a shape cannot afford four sub-tables of the wider window (LeNet layer 5).
struct MsmGeom {
  int W;        // window width (bits)
  int windows;  // signed digits of |s| <= (l-1)/2 < 2^252
  int group;    // local windows per sub-table (Horner length)
  int table;    // multiples 1..2^(W-1) per generator and sub-table
  int sub;      // sub-tables per generator (2 or 4)
};
static inline MsmGeom msm_geom(int W, int sub = kMsmSub) {
  MsmGeom g;
  g.W = W;
  g.windows = 252 / W + 1;
  g.sub = sub;
  g.group = (g.windows + sub - 1) / sub;
  g.table = 1 << (W - 1);
  return g;
}
static inline size_t msm_table_bytes_per_base(int W, int sub = kMsmSub) { return (size_t)sub * ((size_t)1 << (W - 1)) * sizeof(niels_t); }

struct MsmTable {
  niels_t *d_table;   // [geom.sub][n_bases][geom.table]
  size_t n_bases;
  MsmGeom geom;
};
static inline size_t msm_table_entries(size_t n_bases, const MsmGeom &g) { return (size_t)g.sub * n_bases * g.table; }

// bases: n extended points on device. Builds the kMsmSub sub-tables (msm_table_entries(n, g) entries). d_scratch: n points.
void launch_table_build(const ge_t *d_bases, size_t n, const MsmGeom &g, niels_t *d_table, ge_t *d_scratch, cudaStream_t st);

// digits layout: [window][row][col_stride] u16, bit 15 = sign, low bits = magnitude (0 = skip)
static inline size_t msm_col_stride(size_t cols) { return (cols + kMsmColsPerBlock - 1) / kMsmColsPerBlock * kMsmColsPerBlock; }
static inline size_t msm_digits_count(size_t rows, size_t cols, const MsmGeom &g) { return (size_t)g.windows * rows * msm_col_stride(cols); }
// scalars: rows x cols Montgomery elements, row-major with leading dimension ld. extra: optional one more scalar per row
// (the blind, multiplied by base index `cols`), or nullptr. Padding columns get digit 0.
// d_wmask (optional): device word that receives the set of windows holding a non-zero digit (zeroed here first); handed to
// launch_msm_accumulate it lets the kernel skip the other windows without reading their digits.
// d_nonzero (optional): device counter incremented by the number of non-zero digits written (= mixed additions the
// accumulate kernel will execute)
void launch_recode(const fl_t *d_scalars, size_t rows, size_t cols, size_t ld, const fl_t *d_extra, const MsmGeom &g, uint16_t *d_digits,
                   unsigned long long *d_nonzero, cudaStream_t st, uint32_t *d_wmask = nullptr);
// number of column segments the accumulate kernel splits a row into (enough threads to fill 148 SMs)
size_t msm_num_segments(size_t rows, size_t cols_total, const MsmGeom &g);
// partial[(row * geom.group + w') * segs + seg] = sum over the segment's columns and the kMsmSub sub-tables; the optional extra
// column (index cols) uses table base `extra_base`
// blocks_per_sm (0 = as many as fit, six): a cap on the kernel's resident blocks per SM, enforced with a dynamic shared-memory
// request nobody reads. A commitment that runs on a low-priority side stream UNDER latency-bound round kernels (the row half of
// the derefs, prover.cu) must leave registers free on every SM, or those kernels' blocks wait for several 256-us blocks to retire.
void launch_msm_accumulate(const MsmTable &t, const uint16_t *d_digits, size_t rows, size_t cols, bool has_extra, size_t extra_base,
                           size_t segs, ge_t *d_partial, cudaStream_t st, const uint32_t *d_wmask = nullptr, int blocks_per_sm = 0);
// out[row] = sum_w' 2^(W*w') sum_seg partial[row][w'][seg]; d_sums: rows * geom.group scratch points (used when segs > 1);
// d_out (points) and d_comp (32-byte encodings) are optional
void launch_msm_finish(const ge_t *d_partial, size_t rows, size_t segs, const MsmGeom &g, ge_t *d_sums, ge_t *d_out, uint8_t *d_comp,
                       cudaStream_t st);
// first half of the finish only: d_sums[row * geom.group + w'] = sum_seg partial[row][w'][seg] (the caller runs the Horner
// pass itself — the prover's bullet-reduction rounds do it on the host, 60 doublings being far cheaper there than in a
// single GPU thread)
// d_counter / d_seq_word / seq (optional): the last block stores `seq` to *d_seq_word (a host-mapped word) once every sum is
// written — d_sums then normally points into the same mapped slot; d_counter: a zeroed device word, left zero again
void launch_msm_segsum(const ge_t *d_partial, size_t rows, size_t segs, const MsmGeom &g, ge_t *d_sums, cudaStream_t st,
                       unsigned *d_counter = nullptr, uint32_t *d_seq_word = nullptr, uint32_t seq = 0);
// RFC 9496 encoding of n points -> n x 32 bytes
void launch_compress(const ge_t *d_pts, size_t n, uint8_t *d_out, cudaStream_t st);
// decode n x 32 bytes -> points; d_ok[i] = 1 if valid
void launch_decompress(const uint8_t *d_in, size_t n, ge_t *d_pts, uint8_t *d_ok, cudaStream_t st);
// out[i] = a[i] + b[i]
void launch_points_add(const ge_t *a, const ge_t *b, size_t n, ge_t *out, cudaStream_t st);
// generator derivation (Spartan/src/commitments.rs:26-31): n x 64 uniform bytes -> n points
void launch_from_uniform_bytes(const uint8_t *d_in, size_t n, ge_t *d_pts, cudaStream_t st);

}  // namespace vpin
