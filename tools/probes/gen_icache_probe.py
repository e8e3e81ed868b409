#!/usr/bin/env python3
"""Generates tools/probes/icache_probe.cu: a kernel whose code is NBLK blocks of ~1 KB of independent integer instructions behind
a jump table.  Every warp walks the blocks in a pseudo-random order, so the instruction footprint (blocks in use) and the
number of resident warps can be swept independently of everything else: it measures how many instructions per cycle an SM
can FETCH when the hot code does not fit its instruction caches ("What limits it")."""
import sys
NBLK = 256
OPS = 60
out = []
out.append('#include <cstdio>\n#include <cstdlib>\n#include <cuda_runtime.h>\n')
out.append('__global__ void __launch_bounds__(128) k_icache(int nblk, int iters, int spread, unsigned *out)\n{\n')
out.append('    unsigned a0 = threadIdx.x, a1 = blockIdx.x, a2 = 3, a3 = 5, a4 = 7, a5 = 11, a6 = 13, a7 = 17;\n')
out.append('    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;\n')
out.append('    unsigned blk = spread ? (warp * 2654435761u >> 8) % (unsigned)nblk : 0u;\n')
out.append('    for (int it = 0; it < iters; it++)\n    {\n        switch (blk)\n        {\n')
for b in range(NBLK):
    out.append('        case %d:\n' % b)
    for k in range(OPS // 8):
        c = (b * 977 + k * 131 + 12345) & 0xffff
        out.append('            a0 = a0 * 3u + %du; a1 ^= a2 + %du; a2 = (a2 << 1) + a3; a3 += a4 ^ %du; a4 = a4 * 5u + a5; a5 ^= a6 + %du; a6 += a7 << 2; a7 = a7 * 3u + a0;\n' % (c, c + 1, c + 2, c + 3))
    out.append('            break;\n')
out.append('        }\n        blk = (blk * 5u + 1u + (spread ? warp : 0u)) % (unsigned)nblk;\n    }\n')
out.append('    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;\n}\n')
out.append(r'''
int main(int argc, char **argv)
{
    int sms = 148;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); sms = p.multiProcessorCount;
    unsigned *d; cudaMalloc(&d, sizeof(unsigned) * sms * 16 * 128);
    const int iters = 20000;
    const double instr_per_block = argc > 1 ? atof(argv[1]) : 64.0;    // SASS instructions per executed block + loop overhead (from cuobjdump)
    printf("# icache probe: %d SMs, %d iterations per warp, %.0f instructions per block assumed\n", sms, iters, instr_per_block);
    printf("# spread warps_per_sm footprint_blocks ms instr_per_clk_per_sm(at 1.965 GHz)\n");
    for (int spread = 0; spread < 2; spread++)
        for (int wps = 4; wps <= 32; wps *= 2)
            for (int nblk = 2; nblk <= 256; nblk *= 2)
            {
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                const int ctas = sms * wps / 4;
                k_icache<<<ctas, 128>>>(nblk, 200, spread, d);
                cudaEventRecord(e0);
                k_icache<<<ctas, 128>>>(nblk, iters, spread, d);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                const double ipc = (double)wps * iters * instr_per_block / (ms * 1e-3 * 1.965e9);
                printf("%d %2d %3d %8.3f %6.3f\n", spread, wps, nblk, ms, ipc);
                if (cudaGetLastError() != cudaSuccess) { printf("CUDA error\n"); return 1; }
            }
    return 0;
}
''')
open(sys.argv[1] if len(sys.argv) > 1 else 'tools/probes/icache_probe.cu', 'w').write(''.join(out))
