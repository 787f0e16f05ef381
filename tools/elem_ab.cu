// elem_ab.cu -- stand-alone A/B timing of k_element<8,3,3,NeoHookean,force> (the C3 element sweep) for compile-time
// variants of pcx_element.cuh (tuning aid, r2):
//   for v in "2 0" "3 1" "3 2" "4 2"; do set -- $v; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 \
//     -DPCX_ELEM3D_MINB=$1 -DPCX_ELEM3D_XU_SMEM=$2 -Iinclude -Ipancax_b200/csrc -o tools/elem_ab_$1_$2 tools/elem_ab.cu; done
// Hex8 n^3 structured mesh, smooth displacement field, nt = 2 (blockIdx.y), deterministic element vectors; prints the
// average launch time, the energy (must agree between variants) and a checksum of the element vectors.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "pcx_abi.h"
#include "pcx_element.cuh"
using namespace pcx;
#ifndef AB_TANGENT
#define AB_TANGENT false                  // -DAB_TANGENT=true: the stiffness-action sweep K v
#endif
#ifndef AB_MODEL
#define AB_MODEL PCX_MODEL_NEOHOOKEAN     // -DAB_MODEL=2 (Gent) / 3 (Swanson) / 4 (Hencky) / 5 (Blatz-Ko)
#endif

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 128, T = 2, reps = argc > 2 ? atoi(argv[2]) : 10;
  const long long ne = (long long)n * n * n, nn = (long long)(n + 1) * (n + 1) * (n + 1);
  std::vector<double> X(nn * 3), U(T * nn * 3);
  std::vector<int32_t> conn(ne * 8);
  auto nid = [&](int i, int j, int k) { return ((long long)k * (n + 1) + j) * (n + 1) + i; };
  for (int k = 0; k <= n; k++) for (int j = 0; j <= n; j++) for (int i = 0; i <= n; i++) {
    const long long id = nid(i, j, k);
    const double x = (double)i / n, y = (double)j / n, z = (double)k / n;
    X[id * 3] = x; X[id * 3 + 1] = y; X[id * 3 + 2] = z;
    for (int t = 0; t < T; t++) {
      const double s = 0.05 * (t + 1);
      U[(t * nn + id) * 3] = s * sin(2.0 * y) * x; U[(t * nn + id) * 3 + 1] = s * y * (1.0 + 0.3 * z); U[(t * nn + id) * 3 + 2] = -s * 0.2 * z * cos(x);
    }
  }
  for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
    int32_t* c = &conn[(((long long)k * n + j) * n + i) * 8];
    c[0] = nid(i, j, k); c[1] = nid(i + 1, j, k); c[2] = nid(i + 1, j + 1, k); c[3] = nid(i, j + 1, k);
    c[4] = nid(i, j, k + 1); c[5] = nid(i + 1, j, k + 1); c[6] = nid(i + 1, j + 1, k + 1); c[7] = nid(i, j + 1, k + 1);
  }
  ParentTab tab; memset(&tab, 0, sizeof(tab));
  tab.nq = 8; tab.nnpe = 8; tab.nd = 3;
  const double g = 1.0 / sqrt(3.0);
  const int sx[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, sy[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
  for (int q = 0; q < 8; q++) {
    const double xi = sx[q] * g, et = sy[q] * g, ze = sz[q] * g;
    tab.w[q] = 1.0;
    for (int a = 0; a < 8; a++) {
      tab.N[q * TAB_MAX_NNPE + a] = 0.125 * (1 + sx[a] * xi) * (1 + sy[a] * et) * (1 + sz[a] * ze);
      tab.dN[(q * TAB_MAX_NNPE + a) * 3 + 0] = 0.125 * sx[a] * (1 + sy[a] * et) * (1 + sz[a] * ze);
      tab.dN[(q * TAB_MAX_NNPE + a) * 3 + 1] = 0.125 * sy[a] * (1 + sx[a] * xi) * (1 + sz[a] * ze);
      tab.dN[(q * TAB_MAX_NNPE + a) * 3 + 2] = 0.125 * sz[a] * (1 + sx[a] * xi) * (1 + sy[a] * et);
    }
  }
  double *dX, *dU, *dfe, *dep, *dprops; int32_t* dconn; int* dflags;
  const int nblk = (int)((ne + 127) / 128);
  CK(cudaMalloc(&dX, X.size() * 8)); CK(cudaMalloc(&dU, U.size() * 8)); CK(cudaMalloc(&dconn, conn.size() * 4));
  CK(cudaMalloc(&dfe, (size_t)T * ne * 24 * 8)); CK(cudaMalloc(&dep, (size_t)T * nblk * 8)); CK(cudaMalloc(&dflags, 4));
  CK(cudaMalloc(&dprops, 2 * PCX_MAX_PROPS * 8));
  double props[2 * PCX_MAX_PROPS] = {1000.0, 1.0};
  int nprops = 2;
  if (AB_MODEL == PCX_MODEL_GENT) { const double g[3] = {0.833, 0.3846, 3.0}; for (int i = 0; i < 3; i++) props[i] = g[i]; nprops = 3; }
  if (AB_MODEL == PCX_MODEL_SWANSON) { const double g[8] = {10.0, 0.93074321, -0.07673672, 0.0, 0.5, 0.10448312, 1.71691036, 0.01}; for (int i = 0; i < 8; i++) props[i] = g[i]; nprops = 8; }
  if (AB_MODEL == PCX_MODEL_HENCKY) { props[0] = 0.833; props[1] = 0.3846; }
  if (AB_MODEL == PCX_MODEL_BLATZ_KO) { props[0] = 0.3846; nprops = 1; }
  CK(cudaMemcpy(dprops, props, sizeof(props), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dU, U.data(), U.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dconn, conn.data(), conn.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(dflags, 0, 4));
  ElemArgs a; memset(&a, 0, sizeof(a));
  a.ne = ne; a.nn = nn; a.coords = dX; a.conns = dconn; a.us = dU; a.fe = dfe; a.epart = dep; a.flags = dflags;
  a.props_dev = dprops; a.nprops = nprops; a.want_force = 1;
  if (AB_TANGENT) { a.v = dU + nn * 3; a.epart = nullptr; }   // direction: the second step's displacement field (T = 1 below)
  dim3 grid(nblk, AB_TANGENT ? 1 : T);
  constexpr int smem = elem_dyn_smem_bytes(8, 3, 3, AB_TANGENT);
  CK(cudaFuncSetAttribute(k_element<8, 3, 3, AB_MODEL, AB_TANGENT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  auto launch = [&]() { k_element<8, 3, 3, AB_MODEL, AB_TANGENT, 0><<<grid, 128, smem>>>(tab, a); };
  for (int i = 0; i < 3; i++) launch();
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < reps; i++) launch();
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<double> ep((size_t)T * nblk), fe((size_t)1 << 16);
  CK(cudaMemcpy(ep.data(), dep, ep.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(fe.data(), dfe + (size_t)ne * 12, fe.size() * 8, cudaMemcpyDeviceToHost));
  double E = 0, cs = 0; for (double v : ep) E += v; for (size_t i = 0; i < fe.size(); i++) cs += fe[i] * (1 + (i % 7));
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_element<8, 3, 3, AB_MODEL, AB_TANGENT, 0>);
  int flags; cudaMemcpy(&flags, dflags, 4, cudaMemcpyDeviceToHost);
  printf("{\"model\": %d, \"tangent\": %d, \"variant\": \"MINB=%d/%d XU_SMEM=%d/%d ACC_SMEM=%d/%d QUNROLL=%d\", \"n\": %d, \"ms_per_launch\": %.4f, \"regs\": %d, \"local_bytes\": %zu, \"dyn_smem\": %d, \"energy\": %.15e, \"fe_checksum\": %.15e, \"flags\": %d}\n",
         AB_MODEL, (int)AB_TANGENT, PCX_ELEM3D_MINB, PCX_ELEM3D_MINB_T, PCX_ELEM3D_XU_SMEM, PCX_ELEM3D_XU_SMEM_T, PCX_ELEM3D_ACC_SMEM, PCX_ELEM3D_ACC_SMEM_T, PCX_ELEM_QUNROLL, n, ms / reps, fa.numRegs, (size_t)fa.localSizeBytes, smem, E, cs, flags);
  return 0;
}
