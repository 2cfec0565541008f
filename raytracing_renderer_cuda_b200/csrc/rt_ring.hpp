// rt_ring.hpp — counters, entry format and the claim / termination protocol of the barrier-free wavefront kernel
// (k_wf_ring in rt_wavefront.cu, EXPERIMENTAL: RT_WF_GRAIN=ring).  The design is described above that kernel.
//
// The file compiles for the device (nvcc: relaxed GPU-scope loads, atomicAdd, st.release) and for the host (g++:
// __atomic builtins), so that the protocol — the part of the kernel that can lose work, hand it out twice or never
// terminate — is exercised by a multi-threaded simulation on the CPU (tests/ring_sim.cpp, tests/test_ring_protocol.py)
// with exactly the code the kernel runs.  The simulation runs on x86 (total store order): it checks the logic, not
// the memory-model side (release/relaxed scopes), which only a GPU run can.
#pragma once

#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define RT_RING_FN __device__ __forceinline__
#else
#include <sched.h>
#define RT_RING_FN inline
#endif

namespace rtd {
namespace ring {

struct Ring {
    uint32_t* ring;          // [classes][1 << cap_log2] entries: slot (bits 0..23) | tag << 25, tag = 64 | (lap & 63)
    unsigned long long* ctl; // 64-bit counters, 16 words (128 bytes) apart: per class reserve / credits / head, then 2 words
    uint32_t cap_log2;
};
enum : int { RC_RESERVE = 0, RC_CREDITS = 1, RC_HEAD = 2 };

// ---- primitives ----
#if defined(__CUDACC__)
// relaxed, GPU-scope accesses: served by L2, never by a stale L1 line
RT_RING_FN unsigned long long peek(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
RT_RING_FN uint32_t load_entry(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// release store: the thread's earlier writes (its path record) are visible before the entry is (MEMBAR.ALL.GPU + ST in
// SASS; fence.acq_rel / __threadfence() would also invalidate the SM's L1 — CCTL.IVALL — on every push)
#ifndef WF_RING_RELAXED_PUBLISH
RT_RING_FN void publish(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
#else // A/B ONLY (what does the fence cost?): without the release the consumer may read a record that is not there yet
RT_RING_FN void publish(uint32_t* p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
#endif
RT_RING_FN unsigned long long add(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
RT_RING_FN void pause(unsigned ns) { __nanosleep(ns); }
#else
RT_RING_FN unsigned long long peek(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
RT_RING_FN uint32_t load_entry(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
RT_RING_FN void publish(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
RT_RING_FN unsigned long long add(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
RT_RING_FN void pause(unsigned) { sched_yield(); }
#endif

RT_RING_FN unsigned long long* ctl(const Ring& rg, int q, int which) { return rg.ctl + size_t(q * 3 + which) * 16u; }
RT_RING_FN unsigned long long* word(const Ring& rg, int which) { return rg.ctl + size_t(which) * 16u; }
RT_RING_FN uint32_t tag(const Ring& rg, unsigned long long pos) { return 64u | (uint32_t(pos >> rg.cap_log2) & 63u); }
RT_RING_FN uint32_t* entry(const Ring& rg, int q, unsigned long long pos) {
    return rg.ring + (size_t(q) << rg.cap_log2) + (uint32_t(pos) & ((1u << rg.cap_log2) - 1u));
}

struct RingClaim {
    int kind; // < 0: nothing (claim_try: nothing claimable right now; claim_wait: the frame is finished)
    uint32_t n;
    unsigned long long pos, path; // first ring position; first path id (class QNEW only)
};

// NQ classes; entries of class QNEW start new paths and are ignored once `npaths` paths have started
template <int NQ, int QNEW>
struct Protocol {
    enum : int { RC_NEXT_PATH = 3 * NQ, RC_BUSY = 3 * NQ + 1, RC_COUNT = 3 * NQ + 2 };

    // class with the most credits
    static RT_RING_FN int scan(const Ring& rg, unsigned long long npaths, long long& best_c) {
        long long c[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) c[q] = (long long)peek(ctl(rg, q, RC_CREDITS));
        const bool exhausted = peek(word(rg, RC_NEXT_PATH)) >= npaths;
        int best = -1;
        best_c = 0;
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            if (!(q == QNEW && exhausted) && c[q] > best_c) {
                best_c = c[q];
                best = q;
            }
        return best;
    }
    // n credits of class q, then n positions (and n path ids): fails, without side effects that last, when other
    // consumers were faster
    static RT_RING_FN bool take(const Ring& rg, int q, uint32_t n, RingClaim& out) {
        const long long old = (long long)add(ctl(rg, q, RC_CREDITS), (unsigned long long)(-(long long)n));
        if (old < (long long)n) {
            add(ctl(rg, q, RC_CREDITS), (unsigned long long)n);
            return false;
        }
        out.kind = q;
        out.n = n;
        out.pos = add(ctl(rg, q, RC_HEAD), (unsigned long long)n);
        out.path = q == QNEW ? add(word(rg, RC_NEXT_PATH), (unsigned long long)n) : 0ull;
        return true;
    }
    // How many entries a claim asks for when a class holds c credits: a chunk (or what is there), and `batch` chunks at
    // once while the class holds plenty (>= 64 batches) — fewer claims on the counters of the fullest class, which every
    // CTA picks at the same time; never in the tail, where a batch would serialise entries other CTAs could take.
    static RT_RING_FN uint32_t claim_size(long long c, uint32_t chunk, uint32_t batch) {
        if (batch > 1u && c >= 64ll * (long long)(batch * chunk)) return batch * chunk;
        return uint32_t(c < (long long)chunk ? c : (long long)chunk);
    }
    // the caller holds `busy`: one or two attempts, no waiting
    static RT_RING_FN void claim_try(const Ring& rg, unsigned long long npaths, uint32_t chunk, RingClaim& out, uint32_t batch = 1u) {
        out.kind = -1;
        for (int attempt = 0; attempt < 2; ++attempt) {
            long long c;
            const int q = scan(rg, npaths, c);
            if (q < 0) return;
            if (take(rg, q, claim_size(c, chunk, batch), out)) return;
        }
    }
    // waits for work or for the end of the frame; `holding`: the caller still holds `busy` for the chunk it just finished.
    // On success the caller holds `busy` (one count per CTA) until it calls claim_wait(holding = true) again.
    static RT_RING_FN void claim_wait(const Ring& rg, unsigned long long npaths, uint32_t chunk, bool holding, RingClaim& out,
                                      uint32_t batch = 1u) {
        unsigned long long* busy = word(rg, RC_BUSY);
        if (holding) add(busy, ~0ull);
        unsigned backoff = 64u;
        out.kind = -1;
        for (;;) {
            const unsigned long long b0 = peek(busy);
            long long c;
            const int q = scan(rg, npaths, c);
            if (q >= 0) {
                add(busy, 1ull); // before the credits move: a scanner sees either the credits or a busy CTA
                if (take(rg, q, claim_size(c, chunk, batch), out)) return;
                add(busy, ~0ull);
            } else if (b0 == 0ull && peek(busy) == 0ull) {
                return; // nothing queued, nobody who could queue anything
            }
            pause(backoff);
            if (backoff < 2048u) backoff *= 2u;
        }
    }
};

} // namespace ring
} // namespace rtd
