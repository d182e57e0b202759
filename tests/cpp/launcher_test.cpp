// Host-only unit test of the command list (sdrpp_b200/csrc/launcher.h): planning a block makes no CUDA call, per-block records
// land in the descriptor at 16-byte offsets and come back as DEVICE addresses, kernel arguments are packed the way
// cudaLaunchKernel expects them, and two plans of the same block are byte-identical (what the graph cache keys on) while a
// changed grid, argument or event is not. Built and run by tests/test_launcher_host.py (no GPU needed).
#include "../../sdrpp_b200/csrc/launcher.h"
#include <cstdio>
#include <cstdlib>

using namespace sdrpp;

static int fails = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL line %d: %s\n", __LINE__, #c); fails++; } } while (0)

struct Rec { int a; double b; char c; };
static void dummy_kernel() {}

static void plan(Launcher& L, unsigned char* h, unsigned char* d, int grid, int value, cudaEvent_t ev) {
    L.begin_block(h, d, 4096);
    L.deferred = true;
    L.cur_graph = GRAPH_S1;
    const Rec* r = L.push(Rec{ value, 2.5, 'x' });
    L.kernel(SID_MAIN, (const void*)dummy_kernel, dim3(grid), dim3(128), 1024, r, 7, 3.0f);
    L.record(SID_MAIN, ev);
    L.cur_graph = GRAPH_NONE;
    L.wait(SID_TAIL, ev);
    L.cur_graph = GRAPH_TAIL;
    const int* q = L.push(value + 1);
    L.kernel(SID_TAIL, (const void*)dummy_kernel, dim3(2, 3), dim3(64), 0, q);
}

int main() {
    alignas(16) static unsigned char h1[4096], h2[4096];
    unsigned char* d = reinterpret_cast<unsigned char*>(0x7000000000ull);   // never dereferenced
    cudaEvent_t e1 = reinterpret_cast<cudaEvent_t>(0x1234), e2 = reinterpret_cast<cudaEvent_t>(0x5678);
    Launcher A, B;
    plan(A, h1, d, 10, 41, e1);
    plan(B, h2, d, 10, 99, e1);          // other per-block VALUES, same sequence
    CHECK(A.cmds.size() == 4 && B.cmds.size() == 4);
    CHECK(A.kernels == 2);
    // descriptor: records at 16-byte offsets, device addresses returned, host copy holds the values
    CHECK(sizeof(Rec) == 24 && A.desc_used == 32 + sizeof(int));          // second record at the next 16-byte boundary
    CHECK(reinterpret_cast<const Rec*>(h1)->a == 41 && reinterpret_cast<const Rec*>(h2)->a == 99);
    CHECK(*reinterpret_cast<const int*>(h1 + 32) == 42 && *reinterpret_cast<const int*>(h2 + 32) == 100);
    // kernel arguments: pointer (8-byte aligned at 0), int at 8, float at 12
    const Cmd& k = A.cmds[0];
    CHECK(k.type == Cmd::KERNEL && k.nargs == 3 && k.arg_off[0] == 0 && k.arg_off[1] == 8 && k.arg_off[2] == 12);
    const void* p0; std::memcpy(&p0, k.params, 8);
    CHECK(p0 == d);                                                        // the DEVICE address of the record
    int i1; float f2; std::memcpy(&i1, k.params + 8, 4); std::memcpy(&f2, k.params + 12, 4);
    CHECK(i1 == 7 && f2 == 3.0f);
    CHECK(k.grid.x == 10 && k.block.x == 128 && k.smem == 1024 && k.graph == GRAPH_S1 && k.sid == SID_MAIN);
    CHECK(A.cmds[1].type == Cmd::RECORD && A.cmds[1].graph == GRAPH_S1 && A.cmds[2].type == Cmd::WAIT && A.cmds[2].graph == GRAPH_NONE);
    CHECK(A.cmds[3].graph == GRAPH_TAIL && A.cmds[3].grid.y == 3);
    // the same sequence with other per-block values is byte-identical: per-block values live in the descriptor only
    CHECK(std::memcmp(A.cmds.data(), B.cmds.data(), A.cmds.size() * sizeof(Cmd)) == 0);
    // ... and a changed grid, by-value argument or event is not
    Launcher C; plan(C, h2, d, 11, 41, e1);
    CHECK(std::memcmp(A.cmds.data(), C.cmds.data(), sizeof(Cmd)) != 0);
    Launcher D; plan(D, h2, d, 10, 41, e2);
    CHECK(std::memcmp(&A.cmds[1], &D.cmds[1], sizeof(Cmd)) != 0);
    // another descriptor slot is another device address in the arguments: another sequence (one graph per result slot)
    Launcher E; plan(E, h2, d + 8192, 10, 41, e1);
    CHECK(std::memcmp(A.cmds.data(), E.cmds.data(), sizeof(Cmd)) != 0);
    // overflow: push fails softly, kernel() refuses more than 12 arguments
    Launcher F; F.begin_block(h1, d, 16); F.deferred = true;
    CHECK(F.push(Rec{ 1, 1.0, 'a' }) == nullptr);
    CHECK(F.kernel(SID_MAIN, (const void*)dummy_kernel, dim3(1), dim3(1), 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13) == cudaErrorInvalidValue);
    std::printf(fails ? "launcher_test: %d failures\n" : "launcher_test: ok\n", fails);
    return fails ? 1 : 0;
}
