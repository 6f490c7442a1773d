/*
 * launch_count.c -- TEST / BENCH INFRASTRUCTURE ONLY: an LD_PRELOAD shim that counts the kernel launches of the
 * UNMODIFIED reference CUDA program (oracle/_ref/ref_4x4, built from /root/reference/quantum_simulator_4x4.cu with
 * -cudart shared) and prints "launches=<n>" on stderr at exit.  bench.py reports the number next to the
 * reference's own timing line (task 7).  Never loaded by the product.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stddef.h>

typedef struct { unsigned x, y, z; } dim3_t;
typedef int (*launch_fn)(const void *, dim3_t, dim3_t, void **, size_t, void *);
static unsigned long long g_launches;

int cudaLaunchKernel(const void *func, dim3_t grid, dim3_t block, void **args, size_t smem, void *stream)
{
    static launch_fn real;
    if (!real) real = (launch_fn)dlsym(RTLD_NEXT, "cudaLaunchKernel");
    g_launches++;
    return real ? real(func, grid, block, args, smem, stream) : 1;
}

__attribute__((destructor)) static void report(void) { fprintf(stderr, "launches=%llu\n", g_launches); }
