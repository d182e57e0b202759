// Command list of one block and its replay as CUDA graphs.
//
// Every kernel, event and copy of a block goes through a Launcher. In immediate mode (one-shot operations) a command runs
// on the spot. In deferred mode (the front end) the block is first PLANNED into a list of commands -- no CUDA call yet --
// and then executed: the commands tagged for a graph (GRAPH_S1: descriptor upload, ingest, spectrum kernels, fp16 split,
// stage 1; GRAPH_TAIL: the tail and post-detector kernels) are replayed as ONE instantiated CUDA graph each when the very
// same command sequence (kernels, grids, parameters, events) has been seen before, and launched one by one otherwise.
//
// What makes the sequences repeat: everything that changes from block to block (ring positions, decimation phases,
// counts, result-arena pointers) lives in a per-block DESCRIPTOR that the host fills in pinned memory and the graph's
// first node copies to the device; kernels take a pointer into it, so a kernel node's parameters are the same bytes for
// every block that uses the same result slot. An instantiated graph is therefore never updated and never uploaded again:
// one small copy and two graph launches per block cross the host link instead of ten launches with up to 3 KB of
// by-value parameters each (profiles/r1r_*: launches wait for their commands while a host-to-device copy occupies the link).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <functional>
#include <vector>

namespace sdrpp {

enum { SID_MAIN = 0, SID_FFT, SID_TAIL, SID_S1B, SID_D2H, SID_COUNT };
enum { GRAPH_NONE = -1, GRAPH_S1 = 0, GRAPH_TAIL = 1, GRAPH_FFT = 2, GRAPH_KINDS = 3 };

struct Cmd {
    enum Type : int { KERNEL = 0, RECORD, WAIT, MEMCPY, CALL };
    int type, sid, graph, nargs;
    // KERNEL
    const void* fn;
    dim3 grid, block;
    unsigned smem;
    unsigned short arg_off[12];
    unsigned char params[160];
    // RECORD / WAIT
    cudaEvent_t ev;
    // MEMCPY
    void* dst; const void* src; size_t bytes; int kind;
    // CALL: index into Launcher::calls (work that is not expressible as a node; a block that has one is never replayed)
    int call;
};

class Launcher {
public:
    cudaStream_t streams[SID_COUNT] = {};
    bool deferred = false;
    int cur_graph = GRAPH_NONE;
    std::vector<Cmd> cmds;
    std::vector<std::function<cudaError_t(cudaStream_t)>> calls;
    long long kernels = 0;          // kernel commands issued (executed directly or as graph nodes)

    // descriptor arena of the block being planned: host copy (pinned for a front end) and its device twin
    unsigned char* h_desc = nullptr;
    unsigned char* d_desc = nullptr;
    size_t desc_cap = 0, desc_used = 0, desc_flushed = 0;

    void immediate(cudaStream_t st) { for (auto& s : streams) s = st; deferred = false; }
    void begin_block(unsigned char* h, unsigned char* d, size_t cap) {
        cmds.clear(); calls.clear(); h_desc = h; d_desc = d; desc_cap = cap; desc_used = desc_flushed = 0; cur_graph = GRAPH_NONE;
    }

    // Copy a per-block argument record into the descriptor; returns the DEVICE address the kernel reads it from.
    template <class T>
    const T* push(const T& v) {
        const size_t off = (desc_used + 15) & ~(size_t)15;
        if (!h_desc || off + sizeof(T) > desc_cap) return nullptr;
        std::memcpy(h_desc + off, &v, sizeof(T));
        desc_used = off + sizeof(T);
        return reinterpret_cast<const T*>(d_desc + off);
    }

    template <class... A>
    cudaError_t kernel(int sid, const void* fn, dim3 grid, dim3 block, size_t smem, const A&... args) {
        Cmd c;
        std::memset(&c, 0, sizeof(c));
        c.type = Cmd::KERNEL; c.sid = sid; c.graph = cur_graph; c.fn = fn; c.grid = grid; c.block = block; c.smem = (unsigned)smem;
        size_t off = 0;
        bool ok = true;
        auto put = [&](const void* p, size_t size, size_t align) {
            off = (off + align - 1) & ~(align - 1);
            if (off + size > sizeof(c.params) || c.nargs >= 12) { ok = false; return; }
            std::memcpy(c.params + off, p, size);
            c.arg_off[c.nargs++] = (unsigned short)off;
            off += size;
        };
        (put(&args, sizeof(A), alignof(A) > 8 ? 16 : alignof(A)), ...);
        if (!ok) return cudaErrorInvalidValue;
        kernels++;
        return submit(c);
    }
    cudaError_t record(int sid, cudaEvent_t ev) { Cmd c; std::memset(&c, 0, sizeof(c)); c.type = Cmd::RECORD; c.sid = sid; c.graph = cur_graph; c.ev = ev; return submit(c); }
    cudaError_t wait(int sid, cudaEvent_t ev) { Cmd c; std::memset(&c, 0, sizeof(c)); c.type = Cmd::WAIT; c.sid = sid; c.graph = cur_graph; c.ev = ev; return submit(c); }
    cudaError_t memcpy_async(int sid, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
        Cmd c; std::memset(&c, 0, sizeof(c));
        c.type = Cmd::MEMCPY; c.sid = sid; c.graph = cur_graph; c.dst = dst; c.src = src; c.bytes = bytes; c.kind = (int)kind;
        return submit(c);
    }
    // anything else that has to run in stream order on `sid`
    cudaError_t call(int sid, std::function<cudaError_t(cudaStream_t)> f) {
        Cmd c; std::memset(&c, 0, sizeof(c));
        c.type = Cmd::CALL; c.sid = sid; c.graph = cur_graph;
        if (!deferred) return f(streams[sid]);
        c.call = (int)calls.size(); calls.push_back(std::move(f));
        cmds.push_back(c);
        return cudaSuccess;
    }

    // Immediate mode: descriptor records pushed since the last kernel reach the device in front of it (pageable source:
    // cudaMemcpyAsync has staged the bytes when it returns, so the host copy may be reused at once).
    cudaError_t flush_desc(cudaStream_t st) {
        if (desc_used == desc_flushed) return cudaSuccess;
        const cudaError_t e = cudaMemcpyAsync(d_desc + desc_flushed, h_desc + desc_flushed, desc_used - desc_flushed, cudaMemcpyHostToDevice, st);
        desc_flushed = desc_used;
        return e;
    }

    cudaError_t exec(const Cmd& c) {
        cudaStream_t st = streams[c.sid];
        switch (c.type) {
        case Cmd::KERNEL: {
            void* argv[12];
            for (int i = 0; i < c.nargs; i++) argv[i] = const_cast<unsigned char*>(c.params) + c.arg_off[i];
            return cudaLaunchKernel(c.fn, c.grid, c.block, argv, c.smem, st);
        }
        case Cmd::RECORD: return cudaEventRecord(c.ev, st);
        case Cmd::WAIT: return cudaStreamWaitEvent(st, c.ev, 0);
        case Cmd::MEMCPY: return cudaMemcpyAsync(c.dst, c.src, c.bytes, (cudaMemcpyKind)c.kind, st);
        case Cmd::CALL: return calls[(size_t)c.call](st);
        }
        return cudaErrorInvalidValue;
    }

private:
    cudaError_t submit(const Cmd& c) {
        if (deferred) { cmds.push_back(c); return cudaSuccess; }
        if (c.type == Cmd::KERNEL) { if (cudaError_t e = flush_desc(streams[c.sid]); e != cudaSuccess) return e; }
        return exec(c);
    }
};

} // namespace sdrpp
