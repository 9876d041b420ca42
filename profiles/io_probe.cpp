// io_probe.cpp -- host I/O primitives on the GPU box (page cache / tmpfs), to size the CLI's reader
// and writer stages.  g++ -O2 -pthread io_probe.cpp -o io_probe -lcudart
//   ./io_probe /dev/shm 1024      (directory, MiB)
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <class F>
static void par(int nt, size_t n, F f) {
    std::vector<std::thread> th;
    const size_t chunk = ((n + nt - 1) / nt + 4095) & ~(size_t)4095;
    for (int t = 0; t < nt; ++t) {
        const size_t lo = std::min(n, chunk * t), hi = std::min(n, chunk * (t + 1));
        if (hi > lo) th.emplace_back([=] { f(lo, hi - lo); });
    }
    for (auto &x : th) x.join();
}

int main(int argc, char **argv) {
    const std::string dir = argc > 1 ? argv[1] : "/dev/shm";
    const size_t n = (size_t)(argc > 2 ? atoll(argv[2]) : 1024) << 20;
    const std::string src = dir + "/io_probe_src", dst = dir + "/io_probe_dst";
    double t0 = now();
    cudaFree(0);
    printf("cuda context            %.3f s\n", now() - t0);
    char *pin = nullptr;
    t0 = now();
    if (cudaHostAlloc((void **)&pin, n, cudaHostAllocDefault) != cudaSuccess) { printf("cudaHostAlloc failed\n"); return 1; }
    printf("cudaHostAlloc %zu MiB   %.3f s  (%.2f GB/s)\n", n >> 20, now() - t0, n / (now() - t0) / 1e9);
    char *pin2 = nullptr;
    t0 = now();
    void *raw = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_POPULATE, -1, 0);
    printf("mmap populate           %.3f s\n", now() - t0);
    t0 = now();
    if (cudaHostRegister(raw, n, cudaHostRegisterDefault) != cudaSuccess) printf("cudaHostRegister failed\n");
    printf("cudaHostRegister        %.3f s\n", now() - t0);
    pin2 = (char *)raw;
    memset(pin, 'A', n);
    for (size_t i = 99; i < n; i += 100) pin[i] = '\n';
    {   // source file
        int fd = open(src.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
        t0 = now();
        size_t done = 0;
        while (done < n) { ssize_t r = write(fd, pin + done, std::min<size_t>(n - done, 1 << 30)); if (r <= 0) return 2; done += r; }
        printf("write() 1 thread (new)  %.2f GB/s\n", n / (now() - t0) / 1e9);
        close(fd);
    }
    int fd = open(src.c_str(), O_RDONLY);
    for (int rep = 0; rep < 2; ++rep) {
        t0 = now();
        size_t done = 0;
        while (done < n) { ssize_t r = pread(fd, pin2 + done, std::min<size_t>(n - done, 1 << 30), done); if (r <= 0) return 3; done += r; }
        printf("pread 1 thread          %.2f GB/s\n", n / (now() - t0) / 1e9);
    }
    for (int nt : {2, 4, 8, 16}) {
        t0 = now();
        par(nt, n, [&](size_t off, size_t len) {
            size_t d = 0;
            while (d < len) { ssize_t r = pread(fd, pin2 + off + d, len - d, off + d); if (r <= 0) break; d += r; }
        });
        printf("pread %2d threads        %.2f GB/s\n", nt, n / (now() - t0) / 1e9);
    }
    close(fd);
    for (int nt : {1, 4, 8}) {   // pwrite into a fresh file, nt threads
        unlink(dst.c_str());
        int fo = open(dst.c_str(), O_CREAT | O_TRUNC | O_RDWR, 0644);
        t0 = now();
        par(nt, n, [&](size_t off, size_t len) {
            size_t d = 0;
            while (d < len) { ssize_t r = pwrite(fo, pin + off + d, std::min<size_t>(len - d, 64 << 20), off + d); if (r <= 0) break; d += r; }
        });
        printf("pwrite %2d threads (new) %.2f GB/s\n", nt, n / (now() - t0) / 1e9);
        close(fo);
    }
    for (int nt : {1, 4, 8, 16}) {   // ftruncate + mmap + memcpy
        unlink(dst.c_str());
        int fo = open(dst.c_str(), O_CREAT | O_TRUNC | O_RDWR, 0644);
        t0 = now();
        if (ftruncate(fo, n) != 0) return 4;
        char *m = (char *)mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_SHARED, fo, 0);
        if (m == MAP_FAILED) return 5;
        par(nt, n, [&](size_t off, size_t len) { memcpy(m + off, pin + off, len); });
        munmap(m, n);
        printf("mmap+memcpy %2d thr (new) %.2f GB/s\n", nt, n / (now() - t0) / 1e9);
        close(fo);
    }
    {   // overwrite an existing file (pages already allocated)
        int fo = open(dst.c_str(), O_RDWR);
        t0 = now();
        size_t done = 0;
        while (done < n) { ssize_t r = pwrite(fo, pin + done, std::min<size_t>(n - done, 1 << 30), done); if (r <= 0) return 6; done += r; }
        printf("pwrite 1 thread (overwrite) %.2f GB/s\n", n / (now() - t0) / 1e9);
        close(fo);
    }
    for (int nt : {1, 4, 8}) {
        t0 = now();
        par(nt, n, [&](size_t off, size_t len) { memcpy(pin2 + off, pin + off, len); });
        printf("memcpy pinned->pinned %2d thr %.2f GB/s\n", nt, n / (now() - t0) / 1e9);
    }
    unlink(src.c_str());
    unlink(dst.c_str());
    return 0;
}
