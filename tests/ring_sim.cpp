// ring_sim.cpp — multi-threaded CPU simulation of the barrier-free wavefront kernel's queue protocol
// (raytracing_renderer_cuda_b200/csrc/rt_ring.hpp, used by k_wf_ring in rt_wavefront.cu).  Test infrastructure.
//
// Every host thread plays one CTA and runs the kernel's trip — wait for a claim, poll the claimed ring positions for
// the tag of their lap, process the entries, reserve + credit the pushes per class, draw the next claim, publish the
// entries, fall back to the waiting claim — with the protocol functions of rt_ring.hpp themselves.  "Processing" an
// entry is a deterministic function of (path, bounce), so the exact number of entries a frame must process is known.
// Checked: every path starts once and ends once, no slot is ever held twice, a record read after its entry is the
// one written before it, the frame terminates (watchdog), the counters balance, and all of it again over many frames
// on the same ring.  The protocol assumes that a consumer reads the positions it claimed long before the ring has
// advanced by a whole lap (true on a GPU, where resident warps are never descheduled and a lap takes at least two
// trips of every slot; NOT true for preempted host threads).  So: small rings, whose lap tags wrap many times, are run
// with ONE thread; many threads get a ring so large that a lap outlasts any preemption.
//
//   ring_sim <threads> <slots> <cap_log2> <chunk> <paths per frame> <frames> <seed> [batch]   -> prints "ok ..." / exits non-zero
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../raytracing_renderer_cuda_b200/csrc/rt_ring.hpp"

namespace {

constexpr int NQ = 8, QNEW = 0, MAX_DEPTH = 50;
using Ops = rtd::ring::Protocol<NQ, QNEW>;
using rtd::ring::Ring;
using rtd::ring::RingClaim;

struct Record {
    unsigned long long path;
    uint32_t bounce, kind, check;
};

uint64_t mix(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
// where the ray of (path, bounce) goes: a shading class 1..NQ-1, or QNEW = the path ends here
int outcome(uint64_t seed, unsigned long long path, uint32_t bounce) {
    if (bounce >= uint32_t(MAX_DEPTH)) return QNEW;
    const uint64_t h = mix(seed ^ mix(path * 64u + bounce));
    if (h % 100u < 35u) return QNEW;
    return 1 + int((h >> 8) % uint64_t(NQ - 1));
}

struct Sim {
    Ring rg{};
    std::vector<uint32_t> ring_mem;
    std::vector<unsigned long long> ctl_mem;
    std::vector<Record> rec;
    std::vector<std::atomic<int>> held;     // slot currently held by a CTA (1) or queued / unused (0)
    std::vector<std::atomic<int>> started, ended;
    std::atomic<unsigned long long> processed{0};
    std::atomic<int> failed{0};
    uint32_t chunk = 8, batch = 1; // batch: chunks per claim while a class holds plenty (WF_RING_BATCH)
    unsigned long long npaths = 0;
    uint64_t seed = 1;

    void fail(const char* what) {
        if (!failed.exchange(1)) fprintf(stderr, "ring_sim: %s\n", what);
    }

    // k_ring_fill + k_ring_commit (rt_wavefront.cu), single-threaded
    void frame_start(uint32_t n_slots) {
        const unsigned long long r0 = *rtd::ring::ctl(rg, QNEW, rtd::ring::RC_RESERVE);
        for (uint32_t i = 0; i < n_slots; ++i) *rtd::ring::entry(rg, QNEW, r0 + i) = i | (rtd::ring::tag(rg, r0 + i) << 25);
        for (int q = 0; q < NQ; ++q) {
            const unsigned long long r = *rtd::ring::ctl(rg, q, rtd::ring::RC_RESERVE), add = q == QNEW ? n_slots : 0u;
            *rtd::ring::ctl(rg, q, rtd::ring::RC_HEAD) = r;
            *rtd::ring::ctl(rg, q, rtd::ring::RC_RESERVE) = r + add;
            *rtd::ring::ctl(rg, q, rtd::ring::RC_CREDITS) = add;
        }
        *rtd::ring::word(rg, Ops::RC_NEXT_PATH) = 0;
        *rtd::ring::word(rg, Ops::RC_BUSY) = 0;
    }

    // one CTA (k_wf_ring): `chunk` lanes are played one after the other
    void cta() {
        RingClaim cl;
        Ops::claim_wait(rg, npaths, chunk, false, cl, batch);
        std::vector<uint32_t> slot(chunk);
        std::vector<int> out(chunk);
        while (cl.kind >= 0 && !failed.load(std::memory_order_relaxed)) {
            uint32_t count[NQ] = {0};
            const uint32_t n_here = cl.n < chunk ? cl.n : chunk; // a batch claim is worked off a chunk per trip
            for (uint32_t i = 0; i < n_here; ++i) { // every lane polls its position right after the claim, as in the kernel
                const unsigned long long p = cl.pos + i;
                const uint32_t want = rtd::ring::tag(rg, p);
                const uint32_t* e = rtd::ring::entry(rg, cl.kind, p);
                uint32_t v = rtd::ring::load_entry(e);
                while ((v >> 25) != want) {
                    if (failed.load(std::memory_order_relaxed)) return;
                    sched_yield();
                    v = rtd::ring::load_entry(e);
                }
                slot[i] = v & 0xffffffu;
            }
            for (uint32_t i = 0; i < n_here; ++i) {
                const uint32_t s = slot[i];
                out[i] = -1;
                if (s >= rec.size()) {
                    fail("slot index out of range");
                    return;
                }
                if (held[s].exchange(1) != 0) fail("slot handed out twice");
                Record& r = rec[s];
                if (cl.kind == QNEW) {
                    const unsigned long long path = cl.path + i;
                    if (path >= npaths) { // no path left: the slot retires
                        held[s].store(0);
                        continue;
                    }
                    if (started[path].fetch_add(1) != 0) fail("path started twice");
                    r.path = path;
                    r.bounce = 0;
                } else {
                    if (r.kind != uint32_t(cl.kind)) fail("entry in the wrong class (or a stale record)");
                    if (r.check != uint32_t(mix(r.path * 64u + r.bounce))) fail("record does not match its entry");
                    ++r.bounce;
                }
                processed.fetch_add(1, std::memory_order_relaxed);
                const int o = outcome(seed, r.path, r.bounce);
                if (o == QNEW) {
                    if (ended[r.path].fetch_add(1) != 0) fail("path ended twice");
                } else {
                    r.kind = uint32_t(o);
                    r.check = uint32_t(mix(r.path * 64u + r.bounce));
                }
                out[i] = o;
                ++count[o];
            }
            unsigned long long base[NQ];
            for (int q = 0; q < NQ; ++q) {
                base[q] = 0;
                if (count[q]) {
                    base[q] = rtd::ring::add(rtd::ring::ctl(rg, q, rtd::ring::RC_RESERVE), count[q]);
                    rtd::ring::add(rtd::ring::ctl(rg, q, rtd::ring::RC_CREDITS), count[q]);
                }
            }
            RingClaim next;
            if (batch > 1 && cl.n > chunk) { // the rest of a batch claim
                next = cl;
                next.n -= chunk;
                next.pos += chunk;
                if (cl.kind == QNEW) next.path += chunk;
            } else {
                Ops::claim_try(rg, npaths, chunk, next, batch);
            }
            for (uint32_t i = 0; i < n_here; ++i) {
                if (out[i] < 0) continue;
                const unsigned long long p = base[out[i]]++;
                held[slot[i]].store(0);
                rtd::ring::publish(rtd::ring::entry(rg, out[i], p), slot[i] | (rtd::ring::tag(rg, p) << 25));
            }
            cl = next;
            if (cl.kind < 0) Ops::claim_wait(rg, npaths, chunk, true, cl, batch);
        }
    }
};

} // namespace

int main(int argc, char** argv) {
    if (argc < 8) {
        fprintf(stderr, "usage: ring_sim threads slots cap_log2 chunk paths frames seed\n");
        return 64;
    }
    const int threads = atoi(argv[1]);
    const uint32_t slots = uint32_t(atoi(argv[2])), cap_log2 = uint32_t(atoi(argv[3]));
    Sim sim;
    sim.chunk = uint32_t(atoi(argv[4]));
    sim.npaths = strtoull(argv[5], nullptr, 10);
    const int frames = atoi(argv[6]);
    const uint64_t seed0 = strtoull(argv[7], nullptr, 10);
    if (argc > 8) sim.batch = uint32_t(atoi(argv[8]));
    if (slots * 2u > (1u << cap_log2)) {
        fprintf(stderr, "ring_sim: the ring must hold two laps of the slots in use\n");
        return 64;
    }
    sim.ring_mem.assign(size_t(NQ) << cap_log2, 0u);
    sim.ctl_mem.assign(size_t(Ops::RC_COUNT) * 16u, 0ull);
    sim.rg.ring = sim.ring_mem.data();
    sim.rg.ctl = sim.ctl_mem.data();
    sim.rg.cap_log2 = cap_log2;
    sim.rec.resize(slots);
    sim.held = std::vector<std::atomic<int>>(slots);
    sim.started = std::vector<std::atomic<int>>(sim.npaths);
    sim.ended = std::vector<std::atomic<int>>(sim.npaths);

    std::atomic<bool> done{false};
    std::thread watchdog([&] {
        for (int k = 0; k < 1200 && !done.load(); ++k) std::this_thread::sleep_for(std::chrono::milliseconds(100));
        if (!done.load()) {
            fprintf(stderr, "ring_sim: no termination within 120 s (deadlock or lost work)\n");
            _Exit(3);
        }
    });

    unsigned long long total = 0, laps_max = 0;
    for (int f = 0; f < frames; ++f) {
        sim.seed = seed0 + uint64_t(f);
        for (auto& a : sim.held) a.store(0);
        for (auto& a : sim.started) a.store(0);
        for (auto& a : sim.ended) a.store(0);
        sim.processed.store(0);
        const uint32_t n = sim.npaths < slots ? uint32_t(sim.npaths) : slots;
        sim.frame_start(n);
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back([&] { sim.cta(); });
        for (auto& t : pool) t.join();

        unsigned long long expect = 0;
        for (unsigned long long p = 0; p < sim.npaths; ++p) {
            uint32_t b = 0;
            ++expect;
            while (outcome(sim.seed, p, b) != QNEW) {
                ++b;
                ++expect;
            }
            if (sim.started[p].load() != 1 || sim.ended[p].load() != 1) sim.fail("a path did not start and end exactly once");
        }
        if (sim.processed.load() != expect) sim.fail("number of processed entries differs from the paths' bounces");
        for (int q = 0; q < NQ; ++q) {
            const unsigned long long r = *rtd::ring::ctl(sim.rg, q, rtd::ring::RC_RESERVE), h = *rtd::ring::ctl(sim.rg, q, rtd::ring::RC_HEAD);
            const long long c = (long long)*rtd::ring::ctl(sim.rg, q, rtd::ring::RC_CREDITS);
            if (q != QNEW && (r != h || c != 0)) sim.fail("a shading class is not empty at the end of the frame");
            if (c != (long long)(r - h)) sim.fail("credits do not equal reserve - head");
            if ((r >> cap_log2) > laps_max) laps_max = r >> cap_log2;
        }
        if (*rtd::ring::word(sim.rg, Ops::RC_BUSY) != 0) sim.fail("busy counter not zero at the end of the frame");
        for (auto& a : sim.held)
            if (a.load() != 0) sim.fail("a slot is still held at the end of the frame");
        total += expect;
        if (sim.failed.load()) break;
    }
    done.store(true);
    watchdog.join();
    if (sim.failed.load()) return 1;
    printf("ok threads=%d slots=%u cap=2^%u chunk=%u batch=%u paths=%llu frames=%d entries=%llu laps=%llu\n", threads, slots, cap_log2,
           sim.chunk, sim.batch, sim.npaths, frames, total, laps_max);
    return 0;
}
