// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// pcg32.h -- restatement of wjakob/pcg32 @ 70099ead (PCG-XSH-RR 64/32); the reference's ext/pcg32 submodule is
// empty.  Pinned by the pcg-random.org known-answer vector in tests/ (via the oracle's identical restatement).
#pragma once
#include <cstdint>
#include <cstring>

struct pcg32 {
    uint64_t state, inc;
    pcg32() : state(0x853c49e6748fea9bULL), inc(0xda3e39cb94b95bdbULL) { }
    pcg32(uint64_t initstate, uint64_t initseq = 1u) { seed(initstate, initseq); }
    void seed(uint64_t initstate, uint64_t initseq = 1) {
        state = 0U; inc = (initseq << 1u) | 1u; nextUInt(); state += initstate; nextUInt();
    }
    uint32_t nextUInt() {
        uint64_t oldstate = state;
        state = oldstate * 0x5851f42d4c957f2dULL + inc;
        uint32_t xorshifted = (uint32_t) (((oldstate >> 18u) ^ oldstate) >> 27u);
        uint32_t rot = (uint32_t) (oldstate >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    float nextFloat() {
        uint32_t u = (nextUInt() >> 9) | 0x3f800000u; float f; std::memcpy(&f, &u, 4); return f - 1.0f;
    }
};
