// tools/ldlt_probe.cu -- runs dfk::ldlt6_solve (deepfactors_b200/csrc/dfk_gn.cuh) on the host and on the device over the
// same random 6x6 systems (full rank and rank deficient) and prints the largest difference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/variants/ldlt_probe tools/ldlt_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "../deepfactors_b200/csrc/dfk_gn.cuh"

__global__ void k(const float* A, const float* b, float* x, int n)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float M[6][6], bb[6], xx[6];
  for (int r = 0; r < 6; ++r) { bb[r] = b[i * 6 + r]; for (int c = 0; c < 6; ++c) M[r][c] = A[i * 36 + r * 6 + c]; }
  dfk::ldlt6_solve(M, bb, xx);
  for (int r = 0; r < 6; ++r) x[i * 6 + r] = xx[r];
}

// the 80x60 level-2 system of the tracking fixture (1047 -> 1052) at the identity pose: 21 JtJ packed upper, 6 Jtr
static const float kSys[27] = {1044.85437f, 50.6877441f, -196.744049f, -259.680023f, 2850.63501f, -552.098083f, 783.830933f, -189.518951f, -1962.19543f, 218.796036f, 288.450104f, 144.875305f, 554.901672f, -535.422363f, 40.8841782f, 5173.72998f, -902.746765f, -545.784668f, 8176.20068f, -1377.69409f, 854.839539f, -43.661293f, -20.8488197f, 11.5211143f, 57.2899284f, -127.330902f, 10.3622227f};

__global__ void k2(const float* sys, float* pose)
{
  float p[7] = {0, 0, 0, 1, 0, 0, 0};
  float s[27];
  for (int i = 0; i < 27; ++i) s[i] = sys[i];
  dfk::gn_update_pose(s, p);
  for (int i = 0; i < 7; ++i) pose[i] = p[i];
}

int main()
{
  {
    float ph[7] = {0, 0, 0, 1, 0, 0, 0}, pd[7];
    dfk::gn_update_pose(kSys, ph);
    float *ds, *dp;
    cudaMalloc(&ds, sizeof(kSys)); cudaMalloc(&dp, 28);
    cudaMemcpy(ds, kSys, sizeof(kSys), cudaMemcpyHostToDevice);
    k2<<<1, 1>>>(ds, dp);
    cudaMemcpy(pd, dp, 28, cudaMemcpyDeviceToHost);
    printf("gn_update_pose host  : %.6g %.6g %.6g %.6g | %.6g %.6g %.6g\n", ph[0], ph[1], ph[2], ph[3], ph[4], ph[5], ph[6]);
    printf("gn_update_pose device: %.6g %.6g %.6g %.6g | %.6g %.6g %.6g\n", pd[0], pd[1], pd[2], pd[3], pd[4], pd[5], pd[6]);
  }
  const int n = 4096;
  float *A = (float*)malloc(n * 36 * 4), *b = (float*)malloc(n * 24), *xh = (float*)malloc(n * 24), *xd = (float*)malloc(n * 24);
  srand(7);
  for (int i = 0; i < n; ++i) {
    const int rows = (i % 3 == 0) ? 4 : 9;
    float J[9][6];
    for (int r = 0; r < rows; ++r) for (int c = 0; c < 6; ++c) J[r][c] = ((float)rand() / RAND_MAX - 0.5f) * (c < 3 ? 30.f : 300.f);
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) { float s = 0; for (int q = 0; q < rows; ++q) s += J[q][r] * J[q][c]; A[i * 36 + r * 6 + c] = s; }
    float xt[6]; for (int r = 0; r < 6; ++r) xt[r] = (float)rand() / RAND_MAX - 0.5f;
    for (int r = 0; r < 6; ++r) { float s = 0; for (int c = 0; c < 6; ++c) s += A[i * 36 + r * 6 + c] * xt[c]; b[i * 6 + r] = s; }
    float M[6][6], bb[6];
    for (int r = 0; r < 6; ++r) { bb[r] = b[i * 6 + r]; for (int c = 0; c < 6; ++c) M[r][c] = A[i * 36 + r * 6 + c]; }
    dfk::ldlt6_solve(M, bb, xh + i * 6);
  }
  float *dA, *db, *dx;
  cudaMalloc(&dA, n * 36 * 4); cudaMalloc(&db, n * 24); cudaMalloc(&dx, n * 24);
  cudaMemcpy(dA, A, n * 36 * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, b, n * 24, cudaMemcpyHostToDevice);
  k<<<(n + 127) / 128, 128>>>(dA, db, dx, n);
  cudaMemcpy(xd, dx, n * 24, cudaMemcpyDeviceToHost);
  double worst = 0, worst_def = 0;
  for (int i = 0; i < n; ++i) for (int r = 0; r < 6; ++r) {
    const double d = fabs((double)xh[i * 6 + r] - xd[i * 6 + r]);
    if (i % 3 == 0) { if (d > worst_def) worst_def = d; } else if (d > worst) worst = d;
  }
  printf("ldlt6_solve host vs device: max |dx| full rank %.3g, rank deficient %.3g (%s)\n", worst, worst_def, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
