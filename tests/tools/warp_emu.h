// tests/tools/warp_emu.h -- development aid only (never shipped, never used by the product path).
//
// Runs the warp-cooperative kernel code of release-1.0_b200/csrc on the HOST: every lane of the warp is a fiber with its
// own stack, and the collectives of ab200_warp.cuh (ballot, shuffle, syncwarp) are rendezvous points at which the
// scheduler switches to the next lane. Lanes run one after another between two collectives, in an order that can be
// reversed (AB_EMU_ORDER=1) or rotated per rendezvous (AB_EMU_ORDER=2), so a missing wsync() between a store by one lane
// and a load by another shows up as a wrong result instead of passing by luck. Include from exactly one translation
// unit, after defining AB_WARP_EMU and before the kernel headers are used.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <functional>
#include <vector>

namespace ab200 { namespace wemu {

int W = 32;
static const int MAXW = 32;

struct Fiber { void* sp; std::vector<char> stack; bool done; };
static Fiber g_fib[MAXW];
static void* g_main_sp = nullptr;
static int g_cur = -1;
static std::function<void(int)> g_body;
static int g_order = 0;
static unsigned long g_round = 0;

// rendezvous state
static uint64_t g_contrib[MAXW];
static uint64_t g_result[2][MAXW];
static unsigned g_ballot_res[2];
static int g_arrived = 0;
static unsigned long g_gen = 0;
static int g_kind = 0;   // kind of the collective in flight (debug check that all lanes call the same one)
unsigned long g_collectives = 0;

extern "C" void ab_emu_switch(void** save_sp, void* new_sp);
asm(R"(
.text
.globl ab_emu_switch
.type ab_emu_switch,@function
ab_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size ab_emu_switch,.-ab_emu_switch
)");

static void fiber_entry() {
    const int me = g_cur;
    g_body(me);
    g_fib[me].done = true;
    ab_emu_switch(&g_fib[me].sp, g_main_sp);
    abort();
}
static int next_lane(int from) {
    // next unfinished lane after `from` in the current scheduling order
    for (int s = 1; s <= W; ++s) {
        int l = (g_order == 1) ? ((from - s) % W + W) % W : (from + s) % W;
        if (!g_fib[l].done) return l;
    }
    return -1;
}
static bool yield_lane() {
    const int me = g_cur;
    int nx = next_lane(me);
    if (nx < 0 || nx == me) return false;
    g_cur = nx;
    ab_emu_switch(&g_fib[me].sp, g_fib[nx].sp);
    g_cur = me;
    return true;
}
int lane() { return g_cur; }

static uint64_t rendezvous(int kind, uint64_t v, int src) {
    const int me = g_cur;
    if (g_arrived == 0) g_kind = kind;
    else if (g_kind != kind) { fprintf(stderr, "warp_emu: lanes diverged (collective %d vs %d)\n", g_kind, kind); abort(); }
    g_contrib[me] = v;
    const unsigned long my_gen = g_gen;
    if (++g_arrived == W) {
        const int slot = (int)(my_gen & 1);
        unsigned b = 0;
        for (int l = 0; l < W; ++l) { g_result[slot][l] = g_contrib[l]; if (g_contrib[l]) b |= 1u << l; }
        g_ballot_res[slot] = b;
        g_arrived = 0; ++g_gen; ++g_collectives;
    }
    while (g_gen == my_gen) if (!yield_lane()) { fprintf(stderr, "warp_emu: lane %d waits in a collective the other lanes never reach\n", me); abort(); }
    const int slot = (int)(my_gen & 1);
    if (kind == 1) return g_result[slot][src & (MAXW - 1)];
    if (kind == 2) return g_ballot_res[slot];
    return 0;
}
uint64_t shfl64(uint64_t v, int src) { if (src < 0 || src >= W) src = ((src % W) + W) % W; return rendezvous(1, v, src); }
unsigned ballot(int pred) { return (unsigned)rendezvous(2, pred ? 1 : 0, 0); }
void sync() { rendezvous(3, 0, 0); }

// Run body(lane) on `lanes` fibers until all have returned.
void run(int lanes, const std::function<void(int)>& body, size_t stack_bytes = 1 << 20) {
    if (const char* o = getenv("AB_EMU_ORDER")) g_order = atoi(o);
    W = lanes; g_body = body; g_arrived = 0;
    for (int l = 0; l < W; ++l) {
        Fiber& f = g_fib[l];
        f.stack.assign(stack_bytes, 0); f.done = false;
        uintptr_t top = ((uintptr_t)(f.stack.data() + stack_bytes)) & ~(uintptr_t)63;
        void** sp = (void**)top;
        *--sp = nullptr;                    // alignment pad: at entry rsp must be 8 mod 16
        *--sp = (void*)&fiber_entry;        // return address taken by `ret`
        for (int r = 0; r < 6; ++r) *--sp = nullptr;   // rbp rbx r12 r13 r14 r15
        f.sp = (void*)sp;
    }
    for (;;) {
        int first = -1;
        if (g_order == 1) { for (int l = W - 1; l >= 0; --l) if (!g_fib[l].done) { first = l; break; } }
        else { for (int l = 0; l < W; ++l) if (!g_fib[l].done) { first = l; break; } }
        if (first < 0) break;
        g_cur = first;
        ab_emu_switch(&g_main_sp, g_fib[first].sp);
    }
    g_cur = -1;
    if (g_arrived != 0) { fprintf(stderr, "warp_emu: %d lanes stuck in a collective at exit\n", g_arrived); abort(); }
}

} } // namespace ab200::wemu
