// Development harness for the FAST kernels (not part of the product): runs the round-1 kernels
// (k_pyramid + k_fast_r1, kept here as the baseline) and the current k_fast2 pair on the same frames,
// checks that pyramids and corner masks are bit-identical, and times both with CUDA events.
//   nvcc -O3 -std=c++17 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a -o fast_lab fast_lab.cu
//   ./fast_lab frames.raw W H NFRAMES S ITERS
#include "../../ptam_cg_b200/csrc/tracker_kernels.cuh"
#include <vector>
#include <cstdlib>
#include <cstring>
using namespace ptam;
namespace ptam {
// =============================================================================================
// k_fast — FAST-10.  Tile = 128 px x 32 rows per CTA (256 threads), staged in shared memory with a
// 3-row halo by 16-byte loads.  Three phases:
//   1. compass pre-test on packed data: every thread tests 4 adjacent pixels of 4 rows.  Any arc of
//      >= 10 ring pixels contains two ADJACENT compass points (ring 0/4/8/12 = below/right/above/
//      left), so a corner needs (below|above) & (right|left) all brighter than p+t (or all darker
//      than p-t).  The comparisons run two pixels per 32-bit register in 16-bit lanes:
//      bit 15 of  x + (0x8000 - p - t - 1)  is set iff x > p + t,  bit 15 of  (0x8000 + p - t - 1) - x
//      iff x < p - t;  no lane can carry or borrow into its neighbour.
//   2. the surviving candidates (~9 % of level-0 pixels) are compacted into a CTA-wide list, so that
//   3. every thread runs the full 16-pixel ring test (>= 10 contiguous, strict) on one candidate.
// Output: one bit per pixel.  Raster order is restored by k_compact.
// =============================================================================================
constexpr int kFastTW = 128, kFastTH = 32;
constexpr int kFastSW = kFastTW + 32;   // staged bytes per row: x0-16 .. x0+143 (ten 16-byte chunks)
constexpr int kFastSR = kFastTH + 6;    // staged rows: y0-3 .. y0+34


__global__ void __launch_bounds__(256) k_fast_r1(TrackerDev d) {
  __shared__ __align__(16) uint8_t tile[kFastSR * kFastSW];
  __shared__ uint16_t cand_list[kFastTW * kFastTH];
  __shared__ unsigned out_mask[kFastTH][4];
  __shared__ int cand_count;
  const int s = blockIdx.y;
  int l = 0, tb = 0, tiles_x = 1;
  {
    int acc = 0;
#pragma unroll
    for (int k = 0; k < kLevels; k++) {
      const int nx = (d.g.lev[k].w + kFastTW - 1) / kFastTW, ny = (d.g.lev[k].h + kFastTH - 1) / kFastTH;
      if ((int)blockIdx.x >= acc) { l = k; tb = acc; tiles_x = nx; }
      acc += nx * ny;
    }
  }
  const LevelDesc& L = d.g.lev[l];
  const int t = blockIdx.x - tb;
  const int tx = t % tiles_x, ty = t / tiles_x;
  const int x0 = tx * kFastTW, y0 = ty * kFastTH;
  int pitch;
  const uint8_t* im = level_image(d, s, l, pitch);
  const bool al16 = ((pitch & 15) == 0) && ((reinterpret_cast<uintptr_t>(im) & 15) == 0);
  if (threadIdx.x < kFastTH * 4) (&out_mask[0][0])[threadIdx.x] = 0u;
  if (threadIdx.x == 0) cand_count = 0;
  // ---- stage rows y0-3 .. y0+34, bytes x0-16 .. x0+143 (zero outside the image): thread -> one of the
  // ten 16-byte column chunks and rows r0, r0 + 25, so the column tests are done once per thread
  if (threadIdx.x < 25 * (kFastSW / 16)) {
    const int r0 = threadIdx.x / (kFastSW / 16), c = threadIdx.x - r0 * (kFastSW / 16);
    const int x = x0 - 16 + 16 * c;
    const int xmode = (x + 15 < 0 || x >= L.w) ? 0 : ((al16 && x >= 0 && x + 15 < L.w) ? 1 : 2);
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
      const int r = r0 + 25 * rr;
      if (r >= kFastSR) break;
      const int y = y0 - 3 + r;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (xmode && y >= 0 && y < L.h) {
        const uint8_t* p = im + (size_t)y * pitch + x;
        if (xmode == 1) v = __ldg(reinterpret_cast<const uint4*>(p));
        else {
          unsigned w4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
          for (int k = 0; k < 16; k++)
            if (x + k >= 0 && x + k < L.w) w4[k >> 2] |= (unsigned)__ldg(p + k) << (8 * (k & 3));
          v = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
      }
      *reinterpret_cast<uint4*>(&tile[r * kFastSW + 16 * c]) = v;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int thr = d.g.thresholds[l];
  // ---- phase 1: compass pre-test, 4 pixels x 4 rows per thread
  const unsigned K = 0x80008000u - (unsigned)(thr + 1) * 0x00010001u;
  unsigned vmask = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int xx = x0 + 4 * lane + k;
    if (xx >= 3 && xx < L.w - 3) vmask |= 1u << k;
  }
  unsigned cand16 = 0;
#pragma unroll
  for (int rr = 0; rr < 4; rr++) {
    const int ry = 4 * warp + rr;  // row inside the tile; staged row ry+3 is the centre row
    const int y = y0 + ry;
    if (y >= 3 && y < L.h - 3) {  // warp-uniform
      const unsigned* crow = reinterpret_cast<const unsigned*>(&tile[(ry + 3) * kFastSW + 12 + 4 * lane]);
      const unsigned w0 = crow[0], cw = crow[1], w2 = crow[2];
      const unsigned up = *reinterpret_cast<const unsigned*>(&tile[ry * kFastSW + 16 + 4 * lane]);
      const unsigned dn = *reinterpret_cast<const unsigned*>(&tile[(ry + 6) * kFastSW + 16 + 4 * lane]);
      const unsigned lft = __byte_perm(w0, cw, 0x4321);  // pixels x-3 .. x
      const unsigned rgt = __byte_perm(cw, w2, 0x6543);  // pixels x+3 .. x+6
      const unsigned Plo = __byte_perm(cw, 0u, 0x4140), Phi = __byte_perm(cw, 0u, 0x4342);
      const unsigned Blo = K - Plo, Bhi = K - Phi, Dlo = Plo + K, Dhi = Phi + K;
      unsigned b_lo[4], b_hi[4], d_lo[4], d_hi[4];
      const unsigned ring4[4] = {dn, rgt, up, lft};
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const unsigned Xlo = __byte_perm(ring4[q], 0u, 0x4140), Xhi = __byte_perm(ring4[q], 0u, 0x4342);
        b_lo[q] = Xlo + Blo; b_hi[q] = Xhi + Bhi;
        d_lo[q] = Dlo - Xlo; d_hi[q] = Dhi - Xhi;
      }
      const unsigned r_lo = ((b_lo[0] | b_lo[2]) & (b_lo[1] | b_lo[3])) | ((d_lo[0] | d_lo[2]) & (d_lo[1] | d_lo[3]));
      const unsigned r_hi = ((b_hi[0] | b_hi[2]) & (b_hi[1] | b_hi[3])) | ((d_hi[0] | d_hi[2]) & (d_hi[1] | d_hi[3]));
      const unsigned r = ((r_lo >> 15) & 1u) | ((r_lo >> 30) & 2u) | ((r_hi >> 13) & 4u) | ((r_hi >> 28) & 8u);
      cand16 |= (r & vmask) << (4 * rr);
    }
  }
  // ---- phase 2: CTA-wide candidate list (order is irrelevant: the output is a bit mask)
  {
    const int nc = __popc(cand16);
    int inc = nc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += n;
    }
    int base = 0;
    if (lane == 31 && inc > 0) base = atomicAdd(&cand_count, inc);
    base = __shfl_sync(kFull, base, 31);
    int o = base + inc - nc;
    // 16 predicated stores instead of a data-dependent loop: the loop ran for the busiest lane of the warp
    const unsigned e0 = (unsigned)(((4 * warp) << 7) | (4 * lane));
#pragma unroll
    for (int b = 0; b < 16; b++)
      if ((cand16 >> b) & 1u) cand_list[o++] = (uint16_t)(e0 + ((b >> 2) << 7) + (b & 3));
  }
  __syncthreads();
  // ---- phase 3: full ring test, one candidate per thread
  const int n_cand = cand_count;
  for (int ci = threadIdx.x; ci < n_cand; ci += 256) {
    const int e = cand_list[ci];
    const int ry = e >> 7, px = e & 127;
    const uint8_t* c = &tile[(ry + 3) * kFastSW + 16 + px];
    const int p = *c;
    const int hi = p + thr, lo = p - thr;
    unsigned mb = 0, md = 0;
    const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const int v = c[dy[j] * kFastSW + dx[j]];
      mb = __funnelshift_l((unsigned)(hi - v), mb, 1);  // shifts in the sign of (p+t) - v: v > p+t
      md = __funnelshift_l((unsigned)(v - lo), md, 1);  // sign of v - (p-t): v < p-t
    }
    if (run10(mb) || run10(md)) atomicOr(&out_mask[ry][px >> 5], 1u << (px & 31));
  }
  __syncthreads();
  if (threadIdx.x < kFastTH * 4) {
    const int ry = threadIdx.x >> 2, wq = threadIdx.x & 3;
    const int y = y0 + ry, word = (x0 >> 5) + wq;
    if (y < L.h && word < L.nwords) d.mask[(size_t)s * d.g.mask_stride + L.mask_off + (size_t)y * L.nwords + word] = out_mask[ry][wq];
  }
}


}  // namespace ptam

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int kOp>
__global__ void k_ubench(unsigned* out, int iters) {
  unsigned a[8];
  for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 2654435761u + k;
  const unsigned b = out[0] | 0x01020304u;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (kOp == 0) a[k] = __vabsdiffu4(a[k], b);
      if (kOp == 1) a[k] = (a[k] & b) ^ 0x5a5a5a5au;       // LOP3
      if (kOp == 2) a[k] = a[k] * 3u + b;                    // IMAD
      if (kOp == 3) a[k] = __byte_perm(a[k], b, 0x6543);     // PRMT
      if (kOp == 4) { a[k] = __vabsdiffu4(a[k], b); a[k] = (a[k] & b) ^ 0x5a5a5a5au; }  // VABSDIFF4 + LOP3
      if (kOp == 5) { a[k] = __vabsdiffu4(a[k], b); a[k] = a[k] * 3u + b; }             // VABSDIFF4 + IMAD
    }
  }
  unsigned r = 0;
  for (int k = 0; k < 8; k++) r ^= a[k];
  if (r == 0x12345678u) out[1] = r;
}

static void ubench() {
  unsigned* out; CK(cudaMalloc(&out, 64)); CK(cudaMemset(out, 0, 64));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 4096, grid = 148 * 8, block = 256;
  const char* names[6] = {"VABSDIFF4", "LOP3", "IMAD", "PRMT", "VABSDIFF4+LOP3", "VABSDIFF4+IMAD"};
  for (int op = 0; op < 6; op++) {
    float best = 1e9f;
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaEventRecord(e0));
      switch (op) {
        case 0: k_ubench<0><<<grid, block>>>(out, iters); break;
        case 1: k_ubench<1><<<grid, block>>>(out, iters); break;
        case 2: k_ubench<2><<<grid, block>>>(out, iters); break;
        case 3: k_ubench<3><<<grid, block>>>(out, iters); break;
        case 4: k_ubench<4><<<grid, block>>>(out, iters); break;
        case 5: k_ubench<5><<<grid, block>>>(out, iters); break;
      }
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      best = ms < best ? ms : best;
    }
    const double ops = (double)grid * block / 32 * iters * 8 * (op >= 4 ? 2 : 1);
    printf("ubench %-16s %.3f ms  %.1f G warp-inst/s  (%.2f per clk per SM at 1.965 GHz)\n", names[op], best, ops / best * 1e-6, ops / best * 1e-6 / 1.965 / 148);
  }
}

int main(int argc, char** argv) {
  if (argc == 2 && !strcmp(argv[1], "ubench")) { ubench(); return 0; }
  if (argc < 7) { printf("usage: fast_lab frames.raw W H NFRAMES S ITERS\n"); return 2; }
  const int W = atoi(argv[2]), H = atoi(argv[3]), NF = atoi(argv[4]), S = atoi(argv[5]), iters = atoi(argv[6]);
  std::vector<uint8_t> frames((size_t)W * H * NF);
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(frames.data(), 1, frames.size(), f) != frames.size()) { printf("cannot read frames\n"); return 2; }
  fclose(f);
  TrackerDev d{};
  make_geom(d.g, W, H);
  const size_t fbytes = (size_t)W * H;
  const int nsets = 4;  // rotate over > L2 worth of frames
  uint8_t* l0; CK(cudaMalloc(&l0, fbytes * S * nsets));
  for (int k = 0; k < S * nsets; k++) CK(cudaMemcpy(l0 + fbytes * k, frames.data() + fbytes * (k % NF), fbytes, cudaMemcpyHostToDevice));
  uint8_t *pyrA, *pyrB; uint32_t *maskA, *maskB;
  CK(cudaMalloc(&pyrA, d.g.pyr_bytes * S)); CK(cudaMalloc(&pyrB, d.g.pyr_bytes * S));
  CK(cudaMalloc(&maskA, d.g.mask_stride * S * 4)); CK(cudaMalloc(&maskB, d.g.mask_stride * S * 4));
  CK(cudaMemset(pyrA, 0, d.g.pyr_bytes * S)); CK(cudaMemset(pyrB, 0, d.g.pyr_bytes * S));
  CK(cudaMemset(maskA, 0, d.g.mask_stride * S * 4)); CK(cudaMemset(maskB, 0xff, d.g.mask_stride * S * 4));
  d.S = S; d.src.stream_pitch = fbytes; d.src.pitch = W;
  int old_tiles = 0;
  for (int l = 0; l < kLevels; l++) old_tiles += ((d.g.lev[l].w + kFastTW - 1) / kFastTW) * ((d.g.lev[l].h + kFastTH - 1) / kFastTH);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto run_old = [&](int set) {
    TrackerDev a = d; a.src.l0 = l0 + fbytes * S * set; a.pyr = pyrA; a.mask = maskA;
    k_pyramid<<<dim3((W + 63) / 64, (H + 63) / 64, S), 256>>>(a);
    k_fast_r1<<<dim3(old_tiles, S), 256>>>(a);
  };
  const int var = getenv("LAB_VAR") ? atoi(getenv("LAB_VAR")) : 0;
  auto launch_new = [&](const TrackerDev& a, int which) {   // which: 1 level 0, 2 levels 1..3, 3 both
    const dim3 g0(d.g.lev[0].tiles_x, d.g.lev[0].tiles_y, S), g1(d.g.fast_tiles, S);
#define LAB_CASE(V, A, E, M) case V: if (which & 1) k_fast2<true, A, E, M><<<g0, 256>>>(a); if (which & 2) k_fast2<false, A, E, M><<<g1, 256>>>(a); break;
    switch (var) {
      LAB_CASE(0, 1, 1, 6) LAB_CASE(1, 1, 0, 6) LAB_CASE(2, 1, 1, 5) LAB_CASE(3, 1, 0, 5) LAB_CASE(4, 0, 0, 5) LAB_CASE(5, 1, 0, 4) LAB_CASE(6, 1, 0, 7) LAB_CASE(7, 1, 0, 8)
    }
#undef LAB_CASE
  };
  auto run_new = [&](int set) {
    TrackerDev a = d; a.src.l0 = l0 + fbytes * S * set; a.pyr = pyrB; a.mask = maskB;
    launch_new(a, 3);
  };
  if (getenv("LAB_THR")) for (int l = 0; l < kLevels; l++) d.g.thresholds[l] = atoi(getenv("LAB_THR"));
  // ---- correctness on every set
  long bad_pyr = 0, bad_mask = 0, corners = 0;
  std::vector<uint8_t> pa(d.g.pyr_bytes * S), pb(d.g.pyr_bytes * S);
  std::vector<uint32_t> ma(d.g.mask_stride * S), mb(d.g.mask_stride * S);
  for (int set = 0; set < nsets; set++) {
    run_old(set); run_new(set);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(pa.data(), pyrA, pa.size(), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(pb.data(), pyrB, pb.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ma.data(), maskA, ma.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(mb.data(), maskB, mb.size() * 4, cudaMemcpyDeviceToHost));
    for (int s = 0; s < S; s++)
      for (int l = 1; l < kLevels; l++) {
        const LevelDesc& L = d.g.lev[l];
        for (int y = 0; y < L.h; y++)
          bad_pyr += memcmp(&pa[s * d.g.pyr_bytes + L.img_off + (size_t)y * L.pitch], &pb[s * d.g.pyr_bytes + L.img_off + (size_t)y * L.pitch], L.w) != 0;
      }
    for (size_t i = 0; i < ma.size(); i++) { bad_mask += ma[i] != mb[i]; corners += __builtin_popcount(ma[i]); }
  }
  printf("check: pyramid rows differing %ld, mask words differing %ld, corners/frame %.1f\n", bad_pyr, bad_mask, (double)corners / (S * nsets));
  // ---- timing
  for (int rep = 0; rep < 2; rep++) {
    for (int which = 0; which < 2; which++) {
      for (int i = 0; i < 3; i++) which ? run_new(i % nsets) : run_old(i % nsets);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      for (int i = 0; i < iters; i++) which ? run_new(i % nsets) : run_old(i % nsets);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("%s: %.4f ms per %d frames\n", which ? "k_fast2<true>+k_fast2<false>" : "k_pyramid+k_fast_r1       ", ms / iters, S);
    }
  }
  // new kernels separately
  {
    TrackerDev a = d; a.src.l0 = l0; a.pyr = pyrB; a.mask = maskB;
    float ms;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; i++) { a.src.l0 = l0 + fbytes * S * (i % nsets); launch_new(a, 1); }
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("k_fast2<true> alone: %.4f ms\n", ms / iters);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; i++) launch_new(a, 2);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("k_fast2<false> alone: %.4f ms\n", ms / iters);
    TrackerDev b = d; b.src.l0 = l0; b.pyr = pyrA; b.mask = maskA;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; i++) { b.src.l0 = l0 + fbytes * S * (i % nsets); k_fast_r1<<<dim3(old_tiles, S), 256>>>(b); }
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("k_fast_r1 alone: %.4f ms\n", ms / iters);
  }
  return bad_pyr || bad_mask;
}
