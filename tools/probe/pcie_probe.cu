/* pcie_probe.cu — what the host<->device link of the GPU box can do for this engine's copy pattern:
 * pinned-memory D2H / H2D / both at once, in chunks of one 1080p frame (3 133 440 B) and larger, with the pinned buffer
 * allocated from each NUMA node the process may run on (the allocating thread is pinned to the node's CPUs first; Linux
 * places the pages of cudaHostAlloc on the node of the calling CPU).  Prints one JSON object.
 *   nvcc -O2 -o build/pcie_probe tools/probe/pcie_probe.cu ; build/pcie_probe [device]
 */
#define _GNU_SOURCE
#include <cuda_runtime.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static int parse_cpulist(const char *s, cpu_set_t *set)
{
    int n = 0; CPU_ZERO(set);
    while (*s && *s != '\n') {
        int a = (int)strtol(s, (char **)&s, 10), b = a;
        if (*s == '-') b = (int)strtol(s + 1, (char **)&s, 10);
        for (int c = a; c <= b; c++) { CPU_SET(c, set); n++; }
        if (*s == ',') s++;
    }
    return n;
}

static double run(int dir, size_t chunk, size_t total, uint8_t *h, uint8_t *d, uint8_t *h2, uint8_t *d2, cudaStream_t s0, cudaStream_t s1, size_t span)
{
    cudaEvent_t a, b, c; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); CK(cudaEventCreate(&c));
    size_t n = total / chunk;
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a, s0));
    CK(cudaStreamWaitEvent(s1, a, 0));
    for (size_t i = 0; i < n; i++) {
        size_t off = (i * chunk) % (span - chunk + 1); off &= ~(size_t)255;
        if (dir == 0 || dir == 2) CK(cudaMemcpyAsync(h + off, d + off, chunk, cudaMemcpyDeviceToHost, s0));
        if (dir == 1) CK(cudaMemcpyAsync(d + off, h + off, chunk, cudaMemcpyHostToDevice, s0));
        if (dir == 2) CK(cudaMemcpyAsync(d2 + off, h2 + off, chunk, cudaMemcpyHostToDevice, s1));
    }
    CK(cudaEventRecord(c, s1)); CK(cudaStreamWaitEvent(s0, c, 0));
    CK(cudaEventRecord(b, s0));
    CK(cudaEventSynchronize(b));
    float ms = 0; CK(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b); cudaEventDestroy(c);
    return (double)(n * chunk) / ms / 1e6;      /* GB/s per direction */
}

int main(int argc, char **argv)
{
    int dev = argc > 1 ? atoi(argv[1]) : 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    char busid[64]; CK(cudaDeviceGetPCIBusId(busid, sizeof busid, dev));
    for (char *q = busid; *q; q++) if (*q >= 'A' && *q <= 'Z') *q += 32;
    char path[256]; snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", busid);
    int gpu_node = -2; { FILE *f = fopen(path, "r"); if (f) { if (fscanf(f, "%d", &gpu_node) != 1) gpu_node = -2; fclose(f); } }
    cpu_set_t all; sched_getaffinity(0, sizeof all, &all);
    const size_t span = (size_t)1 << 30, total = (size_t)4 << 30;
    uint8_t *d, *d2; CK(cudaMalloc(&d, span)); CK(cudaMalloc(&d2, span));
    cudaStream_t s0, s1; CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    printf("{\"gpu\": \"%s\", \"bus\": \"%s\", \"gpu_numa_node\": %d, \"affinity_cpus\": %d, \"nodes\": [", p.name, busid, gpu_node, CPU_COUNT(&all));
    int first = 1;
    for (int node = 0; node < 16; node++) {
        snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
        FILE *f = fopen(path, "r"); if (!f) break;
        char buf[4096] = ""; if (!fgets(buf, sizeof buf, f)) buf[0] = 0; fclose(f);
        cpu_set_t ns, use; parse_cpulist(buf, &ns); CPU_AND(&use, &ns, &all);
        if (!CPU_COUNT(&use)) { printf("%s{\"node\": %d, \"cpus_allowed\": 0}", first ? "" : ", ", node); first = 0; continue; }
        sched_setaffinity(0, sizeof use, &use);
        uint8_t *h, *h2; CK(cudaHostAlloc(&h, span, cudaHostAllocDefault)); CK(cudaHostAlloc(&h2, span, cudaHostAllocDefault));
        memset(h, 1, span); memset(h2, 2, span);
        printf("%s{\"node\": %d, \"cpus_allowed\": %d", first ? "" : ", ", node, CPU_COUNT(&use)); first = 0;
        const size_t chunks[3] = {3133440, (size_t)16 << 20, (size_t)256 << 20};
        const char *names[3] = {"d2h", "h2d", "both"};
        for (int c = 0; c < 3; c++) for (int dir = 0; dir < 3; dir++) {
            run(dir, chunks[c], total / 4, h, d, h2, d2, s0, s1, span);
            double g = run(dir, chunks[c], total, h, d, h2, d2, s0, s1, span);
            printf(", \"%s_%zuKB_GBps\": %.1f", names[dir], chunks[c] >> 10, g);
        }
        printf("}");
        fflush(stdout);
        cudaFreeHost(h); cudaFreeHost(h2);
        sched_setaffinity(0, sizeof all, &all);
    }
    printf("]}\n");
    return 0;
}
