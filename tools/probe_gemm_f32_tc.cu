// probe_gemm_f32_tc.cu — stand-alone check and timing of mgn_linear_f32_tc (csrc/mgn_gemm_f32_tc.cu) at the shapes the fp32
// MeshGraphNet path has at the c2 size: forward x[M,384] W[128,384]^T + b (ReLU) and the data gradient g_y[M,128] W[128,384]
// (the transposed image: N = 384 in three 128-column launches), against float64 dot products on sampled rows and against the
// library's exact-fp32 SIMT kernels (mgn_linear_fwd / mgn_linear_bwd_data) in the same binary.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 --expt-relaxed-constexpr -o probe_gemm_f32_tc probe_gemm_f32_tc.cu -lcuda
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../modulus_b200/csrc/mgn_misc.cu"
#include "../modulus_b200/csrc/mgn_dense.cu"
#include "../modulus_b200/csrc/mgn_gemm_f32_tc.cu"

#define CK(x)                                                                           \
  do {                                                                                  \
    cudaError_t e_ = (x);                                                               \
    if (e_ != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                          \
    }                                                                                   \
  } while (0)

static float frand() { return (rand() / (float)RAND_MAX) * 2.f - 1.f; }

template <typename F> static float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main(int argc, char** argv) {
  const long long M = argc > 1 ? atoll(argv[1]) : 598502 + 37;  // c2 edge count plus a ragged tail
  const int Kf = 384, Nf = 128;
  srand(99);
  std::vector<float> hx((size_t)M * Kf), hw((size_t)Nf * Kf), hb(Nf), hg((size_t)M * Nf);
  for (auto& v : hx) v = frand();
  for (auto& v : hw) v = frand() * 0.1f;
  for (auto& v : hb) v = frand();
  for (auto& v : hg) v = frand();
  float *dx, *dw, *db, *dg, *dsplit, *dsplit_t, *dout, *dout_ref, *dgx, *dgx_ref;
  int* dstatus;
  CK(cudaMalloc(&dx, hx.size() * 4));
  CK(cudaMalloc(&dw, hw.size() * 4));
  CK(cudaMalloc(&db, hb.size() * 4));
  CK(cudaMalloc(&dg, hg.size() * 4));
  CK(cudaMalloc(&dsplit, 2 * hw.size() * 4));
  CK(cudaMalloc(&dsplit_t, 2 * hw.size() * 4));
  CK(cudaMalloc(&dout, (size_t)M * Nf * 4));
  CK(cudaMalloc(&dout_ref, (size_t)M * Nf * 4));
  CK(cudaMalloc(&dgx, (size_t)M * Kf * 4));
  CK(cudaMalloc(&dgx_ref, (size_t)M * Kf * 4));
  CK(cudaMalloc(&dstatus, 4));
  CK(cudaMemset(dstatus, 0, 4));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dg, hg.data(), hg.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xFF, (size_t)M * Nf * 4));
  CK(cudaMemset(dgx, 0xFF, (size_t)M * Kf * 4));

  int rc = mgn_split_weight_tf32(dw, Nf, Kf, Kf, dsplit, 0, nullptr);
  rc |= mgn_split_weight_tf32(dw, Nf, Kf, Kf, dsplit_t, 1, nullptr);
  printf("split rc=%d\n", rc);
  rc = mgn_linear_f32_tc(dx, Kf, M, Kf, dsplit, Nf, db, MGN_ACT_RELU, dout, Nf, dstatus, nullptr);
  printf("forward rc=%d\n", rc);
  rc = mgn_linear_f32_tc(dg, Nf, M, Nf, dsplit_t, Kf, nullptr, MGN_ACT_NONE, dgx, Kf, dstatus, nullptr);
  printf("dgrad rc=%d\n", rc);
  cudaError_t e = cudaDeviceSynchronize();
  int st = -1;
  if (e == cudaSuccess) CK(cudaMemcpy(&st, dstatus, 4, cudaMemcpyDeviceToHost));
  printf("sync: %s, status word %d\n", cudaGetErrorString(e), st);
  if (e != cudaSuccess) return 3;
  rc = mgn_linear_fwd(MGN_F32, dx, Kf, M, Kf, dw, db, Nf, MGN_ACT_RELU, nullptr, dout_ref, Nf, nullptr);
  printf("SIMT forward rc=%d\n", rc);
  rc = mgn_linear_bwd_data(MGN_F32, dg, M, Nf, dw, Kf, dgx_ref, Kf, nullptr);
  printf("SIMT dgrad rc=%d\n", rc);
  CK(cudaDeviceSynchronize());

  std::vector<float> ho((size_t)M * Nf), hor((size_t)M * Nf), hgx((size_t)M * Kf), hgxr((size_t)M * Kf);
  CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hor.data(), dout_ref, hor.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hgx.data(), dgx, hgx.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hgxr.data(), dgx_ref, hgxr.size() * 4, cudaMemcpyDeviceToHost));
  // float64 dot products on sampled rows (every 1009th row and the last 300 rows: the ragged tile)
  double e_tc = 0, e_simt = 0, eg_tc = 0, eg_simt = 0, sc_f = 0, sc_g = 0;
  long long rows_checked = 0, nanc = 0;
  for (long long m = 0; m < M; ++m) {
    if (!(m % 1009 == 0 || m >= M - 300)) continue;
    ++rows_checked;
    for (int n = 0; n < Nf; ++n) {
      double s = hb[n], sa = 0;
      for (int k = 0; k < Kf; ++k) {
        s += (double)hx[m * Kf + k] * hw[(size_t)n * Kf + k];
        sa += fabs((double)hx[m * Kf + k] * hw[(size_t)n * Kf + k]);
      }
      const double r = s > 0 ? s : 0;
      if (ho[m * Nf + n] != ho[m * Nf + n]) ++nanc;
      e_tc = fmax(e_tc, fabs(ho[m * Nf + n] - r));
      e_simt = fmax(e_simt, fabs(hor[m * Nf + n] - r));
      sc_f = fmax(sc_f, sa);
    }
    for (int k = 0; k < Kf; ++k) {
      double s = 0, sa = 0;
      for (int n = 0; n < Nf; ++n) {
        s += (double)hg[m * Nf + n] * hw[(size_t)n * Kf + k];
        sa += fabs((double)hg[m * Nf + n] * hw[(size_t)n * Kf + k]);
      }
      if (hgx[m * Kf + k] != hgx[m * Kf + k]) ++nanc;
      eg_tc = fmax(eg_tc, fabs(hgx[m * Kf + k] - s));
      eg_simt = fmax(eg_simt, fabs(hgxr[m * Kf + k] - s));
      sc_g = fmax(sc_g, sa);
    }
  }
  // every element against the SIMT result (catches a wrong tile anywhere)
  double d_all = 0, dg_all = 0;
  for (size_t i = 0; i < ho.size(); ++i) d_all = fmax(d_all, fabs((double)ho[i] - hor[i]));
  for (size_t i = 0; i < hgx.size(); ++i) dg_all = fmax(dg_all, fabs((double)hgx[i] - hgxr[i]));
  printf("rows checked against float64: %lld, NaNs %lld\n", rows_checked, nanc);
  printf("forward  M=%lld K=%d N=%d relu: max |err| vs float64  3xTF32 %.3g   SIMT fp32 %.3g   (max sum|a b| %.3g);  max |3xTF32 - SIMT| over ALL elements %.3g\n",
         M, Kf, Nf, e_tc, e_simt, sc_f, d_all);
  printf("dgrad    M=%lld K=%d N=%d     : max |err| vs float64  3xTF32 %.3g   SIMT fp32 %.3g   (max sum|a b| %.3g);  max |3xTF32 - SIMT| over ALL elements %.3g\n",
         M, Nf, Kf, eg_tc, eg_simt, sc_g, dg_all);
  const bool ok = nanc == 0 && st == 0 && e_tc < 1e-6 * sc_f && eg_tc < 1e-6 * sc_g && d_all < 1e-4 && dg_all < 1e-4;
  printf("%s\n", ok ? "PASS" : "FAIL");

  const float t_tc = time_ms([&] { mgn_linear_f32_tc(dx, Kf, M, Kf, dsplit, Nf, db, MGN_ACT_RELU, dout, Nf, dstatus, nullptr); }, 10);
  const float t_simt = time_ms([&] { mgn_linear_fwd(MGN_F32, dx, Kf, M, Kf, dw, db, Nf, MGN_ACT_RELU, nullptr, dout_ref, Nf, nullptr); }, 5);
  const float tg_tc = time_ms([&] { mgn_linear_f32_tc(dg, Nf, M, Nf, dsplit_t, Kf, nullptr, MGN_ACT_NONE, dgx, Kf, dstatus, nullptr); }, 10);
  const float tg_simt = time_ms([&] { mgn_linear_bwd_data(MGN_F32, dg, M, Nf, dw, Kf, dgx_ref, Kf, nullptr); }, 5);
  const double fl = 2.0 * M * Kf * Nf;
  const double by_f = (double)M * (Kf + Nf) * 4, by_g = (double)M * (Nf + Kf) * 4;
  printf("forward : 3xTF32 %.3f ms (%.1f TFLOP/s useful, %.0f GB/s algorithmic)   SIMT fp32 %.3f ms   speed-up %.1f x\n", t_tc,
         fl / t_tc * 1e-9, by_f / t_tc * 1e-6, t_simt, t_simt / t_tc);
  printf("dgrad   : 3xTF32 %.3f ms (%.1f TFLOP/s useful, %.0f GB/s algorithmic)   SIMT fp32 %.3f ms   speed-up %.1f x\n", tg_tc,
         fl / tg_tc * 1e-9, by_g / tg_tc * 1e-6, tg_simt, tg_simt / tg_tc);
  CK(cudaMemcpy(&st, dstatus, 4, cudaMemcpyDeviceToHost));
  printf("status word after timing: %d\n", st);
  return ok ? 0 : 1;
}
