// host-side check of block_exchange_sort against the scalar exchange sort (runs on a GPU)
#include <cstdio>
#include <vector>
#include <cstdlib>
#include "../fccf_pcr_b200/csrc/fccf_dev.cuh"
using namespace fccf;
template <typename K> __global__ void k(K* key, int* perm, int n) { __shared__ unsigned long long s[40]; block_exchange_sort(key, perm, n, s); }
template <typename K> void ref(std::vector<K>& k, std::vector<int>& p) { int n = k.size(); for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) if (k[i] < k[j]) { std::swap(k[i], k[j]); std::swap(p[i], p[j]); } }
int main() {
  int bad = 0;
  for (int trial = 0; trial < 100; trial++) {
    int n = trial < 5 ? trial : (rand() % 3000 + 1);
    int nt = (trial % 3 == 0) ? 1024 : ((trial % 3 == 1) ? 256 : 96);
    std::vector<int> key(n), perm(n);
    int mode = trial % 5;
    for (int i = 0; i < n; i++) { key[i] = mode == 0 ? 1 + (rand() % 100 == 0) * (rand() % 50) : (mode == 1 ? rand() % 5 : (mode == 2 ? rand() : (mode == 3 ? i : (rand() % 40 == 0 ? 0 : 1 + (rand() % 6 == 0) * (rand() % 20))))); perm[i] = i; }
    std::vector<int> rk = key, rp = perm; ref(rk, rp);
    int *dk, *dp; cudaMalloc(&dk, 4 * (n + 1)); cudaMalloc(&dp, 4 * (n + 1));
    cudaMemcpy(dk, key.data(), 4 * n, cudaMemcpyHostToDevice); cudaMemcpy(dp, perm.data(), 4 * n, cudaMemcpyHostToDevice);
    k<int><<<1, nt>>>(dk, dp, n);
    cudaMemcpy(key.data(), dk, 4 * n, cudaMemcpyDeviceToHost); cudaMemcpy(perm.data(), dp, 4 * n, cudaMemcpyDeviceToHost);
    bool ok = key == rk && perm == rp;
    if (!ok) { bad++; printf("int trial %d n %d nt %d MISMATCH\n", trial, n, nt); }
    cudaFree(dk); cudaFree(dp);
    // floats
    std::vector<float> fk(n), frk; std::vector<int> fp(n), frp;
    for (int i = 0; i < n; i++) { fk[i] = mode == 1 ? (rand() % 7) * 0.25f - 0.5f : (float)rand() / RAND_MAX; fp[i] = i; }
    frk = fk; frp = fp; ref(frk, frp);
    float* dfk; cudaMalloc(&dfk, 4 * (n + 1)); cudaMalloc(&dp, 4 * (n + 1));
    cudaMemcpy(dfk, fk.data(), 4 * n, cudaMemcpyHostToDevice); cudaMemcpy(dp, fp.data(), 4 * n, cudaMemcpyHostToDevice);
    k<float><<<1, nt>>>(dfk, dp, n);
    cudaMemcpy(fk.data(), dfk, 4 * n, cudaMemcpyDeviceToHost); cudaMemcpy(fp.data(), dp, 4 * n, cudaMemcpyDeviceToHost);
    if (!(fk == frk && fp == frp)) { bad++; printf("float trial %d n %d nt %d MISMATCH\n", trial, n, nt); }
    cudaFree(dfk); cudaFree(dp);
  }
  printf("block_exchange_sort: %d mismatches, last error %s\n", bad, cudaGetErrorString(cudaGetLastError()));
  return bad != 0;
}
