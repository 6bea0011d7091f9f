// Standalone harness for the dense S1 homography warp through the C ABI (no Python, no torch):
// cfg2 / cfg4 shaped synthetic pairs, both directions in one fused launch, CUDA-event timing with an L2
// flush between iterations, optional per-CTA timeline dump (library built with -DDMH_TILE_DEBUG).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tile_bench tools/tile_bench.cu \
//        -Ldmhomo_b200 -ldmhomo -Xlinker -rpath -Xlinker '$ORIGIN/../dmhomo_b200'
//   tools/tile_bench [B C h w rho iters fwd_only [key=value ...]]     (key=value: dmh_set_tuning knobs, e.g. tile=0)
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/dmhomo.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

extern "C" int dmh_tile_debug_dump(void* host, int bytes) __attribute__((weak));

static unsigned long long rng = 0x9E3779B97F4A7C15ull;
static float urand() {
  rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
  return (float)((rng >> 40) & 0xFFFFFF) / 16777216.0f;
}

__global__ void fill_noise(float* p, size_t n, unsigned seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned v = (unsigned)i * 2654435761u ^ seed;
    v ^= v >> 15; v *= 2246822519u; v ^= v >> 13; v *= 3266489917u; v ^= v >> 16;
    p[i] = (v >> 8) * (1.0f / 16777216.0f);
  }
}

int main(int argc, char** argv) {
  int B = 64, C = 1, h = 320, w = 576, iters = 20, fwd_only = 0;
  float rho = 32.f;
  if (argc > 4) { B = atoi(argv[1]); C = atoi(argv[2]); h = atoi(argv[3]); w = atoi(argv[4]); }
  if (argc > 5) rho = (float)atof(argv[5]);
  if (argc > 6) iters = atoi(argv[6]);
  if (argc > 7) fwd_only = atoi(argv[7]);
  for (int i = 8; i < argc; ++i) {   // key=value development knobs (dmh_set_tuning)
    char* eq = strchr(argv[i], '=');
    if (!eq) continue;
    *eq = 0;
    if (dmh_set_tuning(argv[i], atoi(eq + 1))) { printf("tuning: %s\n", dmh_last_error_string()); return 1; }
  }
  const size_t plane = (size_t)h * w, img = (size_t)B * C * plane;

  float *i1, *i2, *g1, *g2, *o1, *o2, *Hf, *Hb, *gHf, *gHb, *src, *dst, *flush;
  uint8_t *v1, *v2;
  double* acc;
  CK(cudaMalloc(&i1, img * 4)); CK(cudaMalloc(&i2, img * 4));
  CK(cudaMalloc(&g1, img * 4)); CK(cudaMalloc(&g2, img * 4));
  CK(cudaMalloc(&o1, img * 4)); CK(cudaMalloc(&o2, img * 4));
  CK(cudaMalloc(&v1, (size_t)B * plane)); CK(cudaMalloc(&v2, (size_t)B * plane));
  CK(cudaMalloc(&Hf, B * 36)); CK(cudaMalloc(&Hb, B * 36)); CK(cudaMalloc(&gHf, B * 36)); CK(cudaMalloc(&gHb, B * 36));
  CK(cudaMalloc(&src, B * 32)); CK(cudaMalloc(&dst, B * 32)); CK(cudaMalloc(&acc, 2 * B * 8));
  const size_t flush_n = 64u << 20;   // 256 MB > 126 MB L2
  CK(cudaMalloc(&flush, flush_n * 4));
  fill_noise<<<1024, 256>>>(i1, img, 1u);
  fill_noise<<<1024, 256>>>(i2, img, 2u);

  // random 4-point homographies through the library's own DLT
  std::vector<float> hs(B * 8), hd(B * 8);
  for (int dir = 0; dir < 2; ++dir) {
    for (int b = 0; b < B; ++b) {
      const float cx[4] = {0.f, (float)(w - 1), 0.f, (float)(w - 1)}, cy[4] = {0.f, 0.f, (float)(h - 1), (float)(h - 1)};
      for (int k = 0; k < 4; ++k) {
        hs[b * 8 + 2 * k] = cx[k]; hs[b * 8 + 2 * k + 1] = cy[k];
        hd[b * 8 + 2 * k] = cx[k] + (urand() * 2 - 1) * rho; hd[b * 8 + 2 * k + 1] = cy[k] + (urand() * 2 - 1) * rho;
      }
    }
    CK(cudaMemcpy(src, hs.data(), B * 32, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dst, hd.data(), B * 32, cudaMemcpyHostToDevice));
    if (dmh_dlt4_forward(src, dst, dir ? Hb : Hf, B, nullptr)) { printf("dlt: %s\n", dmh_last_error_string()); return 1; }
  }
  CK(cudaDeviceSynchronize());

  dmh_warp_desc d[2];
  memset(d, 0, sizeof(d));
  for (int t = 0; t < 2; ++t) {
    d[t].struct_size = sizeof(dmh_warp_desc);
    d[t].sampler = DMH_S1; d[t].param_kind = DMH_PARAM_HOMOGRAPHY;
    d[t].B = B; d[t].C = C; d[t].Hs = h; d[t].Ws = w; d[t].h = h; d[t].w = w; d[t].divide = 1;
    d[t].src = t ? i1 : i2; d[t].param = t ? Hb : Hf;
    if (fwd_only) {
      d[t].loss_form = DMH_LOSS_NONE;
      d[t].out = t ? o1 : o2; d[t].valid = t ? v1 : v2;
    } else {
      d[t].loss_form = DMH_LOSS_MASKED_DIFF; d[t].use_border_mask = 1; d[t].compute_grads = 1;
      d[t].grad_loss_scale = 1.0f / (float)img;
      d[t].target = t ? i2 : i1;
      d[t].loss_acc = acc + t * B;
      d[t].grad_src = t ? g1 : g2; d[t].grad_target = t ? g2 : g1; d[t].grad_param = t ? gHb : gHf;
    }
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9f, sum = 0.f;
  for (int it = 0; it < iters + 3; ++it) {
    CK(cudaMemsetAsync(g1, 0, img * 4)); CK(cudaMemsetAsync(g2, 0, img * 4));
    CK(cudaMemsetAsync(acc, 0, 2 * B * 8)); CK(cudaMemsetAsync(gHf, 0, B * 36)); CK(cudaMemsetAsync(gHb, 0, B * 36));
    fill_noise<<<2048, 256>>>(flush, flush_n, (unsigned)it);   // evict the inputs from L2
    CK(cudaEventRecord(e0));
    const int rc = dmh_warp_forward(d, 2, nullptr);
    CK(cudaEventRecord(e1));
    if (rc) { printf("warp: %s\n", dmh_last_error_string()); return 1; }
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (it >= 3) { best = fminf(best, ms); sum += ms; }
  }
  const double px = 2.0 * B * plane;
  const double bytes = px * (fwd_only ? (8.0 * C + 1) : (24.0 * C + 1));
  std::vector<double> hacc(2 * B);
  CK(cudaMemcpy(hacc.data(), acc, 2 * B * 8, cudaMemcpyDeviceToHost));
  double l = 0; for (double v : hacc) l += v;
  printf("B=%d C=%d %dx%d %s: best %.4f ms  mean %.4f ms  %.1f Gpix/s  %.0f GB/s algorithmic (best)  loss-sum %.6f\n", B, C, h, w,
         fwd_only ? "forward" : "fused", best, sum / iters, px / best * 1e-6, bytes / best * 1e-6, l / (double)img);

  if (dmh_tile_debug_dump) {
    // per CTA: smid, start ns, end ns, tiles, spins
    std::vector<unsigned long long> dbg(1024 * 10);
    const int n = dmh_tile_debug_dump(dbg.data(), (int)(dbg.size() * 8));
    unsigned long long t0 = ~0ull, t1 = 0;
    for (int i = 0; i < n; ++i) { if (dbg[i * 10 + 1] < t0) t0 = dbg[i * 10 + 1]; if (dbg[i * 10 + 2] > t1) t1 = dbg[i * 10 + 2]; }
    printf("debug: %d CTAs, kernel span %.1f us\n", n, (t1 - t0) * 1e-3);
    {   // per-CTA statistics (min / mean / max): end time, producer phases, consumer-warp-0 wait
      const char* names[6] = {"end_us", "loader done-wait kcycles", "loader free-wait (shared buffers)", "loader queue-wait", "(unused)", "consumer-warp-0 full-wait kcycles"};
      for (int f = 0; f < 6; ++f) {
        double mn = 1e30, mx = -1e30, sm = 0;
        for (int i = 0; i < n; ++i) {
          const double v = (f == 0) ? (dbg[i * 10 + 2] - t0) * 1e-3 : dbg[i * 10 + 4 + f] / 1000.0;
          mn = v < mn ? v : mn; mx = v > mx ? v : mx; sm += v;
        }
        printf("  %-36s min %8.1f  mean %8.1f  max %8.1f\n", names[f], mn, sm / (n > 0 ? n : 1), mx);
      }
    }
    for (int i = 0; i < n; ++i)
      printf("cta %3d sm %3llu start %7.1f end %7.1f us tiles %llu spins %llu  kcycles: done-wait %llu drain %llu claim %llu stage %llu full-wait %llu\n",
             i, dbg[i * 10], (dbg[i * 10 + 1] - t0) * 1e-3, (dbg[i * 10 + 2] - t0) * 1e-3, dbg[i * 10 + 3], dbg[i * 10 + 4],
             dbg[i * 10 + 5] / 1000, dbg[i * 10 + 6] / 1000, dbg[i * 10 + 7] / 1000, dbg[i * 10 + 8] / 1000, dbg[i * 10 + 9] / 1000);
  }
  return 0;
}
