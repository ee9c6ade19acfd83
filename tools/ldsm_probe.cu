// ldsm_probe.cu -- prints the thread/register/byte mapping of
// ldmatrix.sync.aligned.m16n16.{x1,x2}.trans.shared.b8 on sm_100a (LDSM.8.MT1616).
// smem byte (matrix m, row r, col c) holds the value m*256 + r*16 + c ... packed as (m<<8 | r<<4 | c) & 0xFF
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__global__ void probe(uint32_t* out) {
  __shared__ __align__(128) uint8_t sm[2][16][16];
  for (int i = threadIdx.x; i < 512; i += 32) {
    int m = i >> 8, r = (i >> 4) & 15, c = i & 15;
    sm[m][r][c] = (uint8_t)((m << 7) | (r << 3 & 0x78) | (c & 7));  // lossy tag, see second pass
  }
  __syncwarp();
  // pass A: tag = row index (0..15) in every byte; pass B: tag = column index
  for (int pass = 0; pass < 2; pass++) {
    for (int i = threadIdx.x; i < 512; i += 32) {
      int m = i >> 8, r = (i >> 4) & 15, c = i & 15;
      sm[m][r][c] = (uint8_t)((m << 4) | (pass == 0 ? r : c));
    }
    __syncwarp();
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(&sm[threadIdx.x >> 4][threadIdx.x & 15][0]);
    uint32_t a, b, c, d;
    asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0, %1, %2, %3}, [%4];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
    out[(pass * 32 + threadIdx.x) * 4 + 0] = a;
    out[(pass * 32 + threadIdx.x) * 4 + 1] = b;
    out[(pass * 32 + threadIdx.x) * 4 + 2] = c;
    out[(pass * 32 + threadIdx.x) * 4 + 3] = d;
    __syncwarp();
  }
}
int main() {
  uint32_t* d; cudaMalloc(&d, 4 * 64 * 4);
  probe<<<1, 32>>>(d);
  uint32_t h[256]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("lane: reg0..3, each byte as (matrix,row,col) of the source\n");
  for (int l = 0; l < 32; l++) {
    printf("lane %2d:", l);
    for (int r = 0; r < 4; r++) {
      printf(" [");
      for (int b = 0; b < 4; b++) {
        int vr = (h[l * 4 + r] >> (8 * b)) & 0xFF, vc = (h[(32 + l) * 4 + r] >> (8 * b)) & 0xFF;
        printf("%d:%2d,%2d%s", vr >> 4, vr & 15, vc & 15, b < 3 ? " " : "");
      }
      printf("]");
    }
    printf("\n");
  }
  return 0;
}
