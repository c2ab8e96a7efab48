// Probe: do two kernels on two streams share the SMs on this box?  (one-off diagnostic for the stream-overlap experiments of
// DESIGN.md section 4; not part of the library)   nvcc -arch=sm_100a -o build/overlap_probe tests/stream_overlap_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void spin(long long cycles, int* sink) {
    extern __shared__ char sm[];
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (sink && threadIdx.x == 0 && cycles < 0) sink[blockIdx.x] = sm[0];
}

static float run(cudaStream_t a, cudaStream_t b, int ctas, int thrA, size_t smA, int thrB, size_t smB, bool second) {
    cudaEvent_t e0, e1, fork, join;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventCreateWithFlags(&fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&join, cudaEventDisableTiming);
    const long long cyc = 400000;          // ~0.2 ms at 1.9 GHz
    cudaDeviceSynchronize();
    cudaEventRecord(e0, a);
    cudaEventRecord(fork, a);
    if (second) {
        cudaStreamWaitEvent(b, fork, 0);
        spin<<<ctas, thrB, smB, b>>>(cyc, nullptr);
        cudaEventRecord(join, b);
    }
    spin<<<ctas, thrA, smA, a>>>(cyc, nullptr);
    if (second) cudaStreamWaitEvent(a, join, 0);
    cudaEventRecord(e1, a);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

__global__ void tiny(int* sink) { if (sink) sink[0] = 1; }

// main stream: A (120 KB smem), a single-warp kernel WITHOUT shared memory, A again; side stream: B (8 KB smem), forked before the
// first A and joined after the second.  Does the small kernel's different L1 / shared-memory split drain the side kernel?
static float run3(cudaStream_t a, cudaStream_t b, bool side, bool tiny_max_shared) {
    cudaEvent_t e0, e1, fork, join;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventCreateWithFlags(&fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&join, cudaEventDisableTiming);
    cudaFuncSetAttribute(tiny, cudaFuncAttributePreferredSharedMemoryCarveout, tiny_max_shared ? cudaSharedmemCarveoutMaxShared : cudaSharedmemCarveoutDefault);
    const long long cyc = 400000;
    cudaDeviceSynchronize();
    cudaEventRecord(e0, a);
    cudaEventRecord(fork, a);
    if (side) {
        cudaStreamWaitEvent(b, fork, 0);
        spin<<<148, 256, 8 * 1024, b>>>(2 * cyc, nullptr);          // ~0.42 ms
        cudaEventRecord(join, b);
    }
    spin<<<148, 512, 120 * 1024, a>>>(cyc / 2, nullptr);           // ~0.1 ms
    tiny<<<1, 32, 0, a>>>(nullptr);
    spin<<<148, 512, 120 * 1024, a>>>(cyc / 2, nullptr);           // ~0.1 ms
    if (side) cudaStreamWaitEvent(a, join, 0);
    cudaEventRecord(e1, a);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    int lo, hi;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    cudaStream_t own, side_hi, side_lo, side_blocking;
    cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking);
    cudaStreamCreateWithPriority(&side_hi, cudaStreamNonBlocking, hi);
    cudaStreamCreateWithPriority(&side_lo, cudaStreamNonBlocking, lo);
    cudaStreamCreate(&side_blocking);
    cudaFuncSetAttribute(spin, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct { const char* name; cudaStream_t a, b; } pairs[] = {
        {"main = legacy default, side = non-blocking high priority", 0, side_hi},
        {"main = legacy default, side = non-blocking low priority", 0, side_lo},
        {"main = own non-blocking, side = non-blocking high priority", own, side_hi},
        {"main = own non-blocking, side = non-blocking low priority", own, side_lo},
        {"main = own non-blocking, side = blocking", own, side_blocking},
    };
    struct { const char* name; int thrA; size_t smA; int thrB; size_t smB; } shapes[] = {
        {"A 512 thr / 120 KB smem, B 256 thr / no smem", 512, 120 * 1024, 256, 0},
        {"A 512 thr / 120 KB smem, B 256 thr / 8 KB smem", 512, 120 * 1024, 256, 8 * 1024},
        {"A 512 thr / no smem,     B 256 thr / no smem", 512, 0, 256, 0},
    };
    for (auto& sh : shapes) {
        printf("%s (148 CTAs each, each kernel ~0.21 ms alone)\n", sh.name);
        for (auto& p : pairs) {
            run(p.a, p.b, 148, sh.thrA, sh.smA, sh.thrB, sh.smB, true);
            const float alone = run(p.a, p.b, 148, sh.thrA, sh.smA, sh.thrB, sh.smB, false);
            const float both = run(p.a, p.b, 148, sh.thrA, sh.smA, sh.thrB, sh.smB, true);
            printf("    %-62s alone %.3f ms, with the side kernel %.3f ms -> %s\n", p.name, alone, both, both < 1.5 * alone ? "concurrent" : "SERIAL");
        }
    }
    printf("main: A(120 KB) , single-warp kernel without shared memory , A(120 KB) = ~0.21 ms; side: B(8 KB) ~0.42 ms alone\n");
    for (int mx = 0; mx < 2; mx++) {
        run3(own, side_hi, true, mx);
        const float alone = run3(own, side_hi, false, mx), both = run3(own, side_hi, true, mx);
        printf("    small kernel carve-out %-10s main alone %.3f ms, with the side kernel %.3f ms (concurrent = 0.42, serial = 0.63)\n",
               mx ? "max shared" : "default", alone, both);
    }
    return 0;
}
