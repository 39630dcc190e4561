// Microbenchmark and self-check of the host arithmetic coder (phase 2 of the entropy stage).
//   g++ -std=c++17 -O2 -I ../include -I ../cairo_b200/csrc/host -o /tmp/abac_bench abac_bench.cpp -lpthread
//   /tmp/abac_bench [p0]        time per bin on 173k random bins with P(bin = 0) = p0
//   /tmp/abac_bench check       fast coder vs a bit-at-a-time coder written straight from the
//                               algorithm (division, one renormalisation step per loop turn) on many
//                               bin distributions and lengths, and the reciprocal identity behind
//                               the fast split point for every tot < 2^23 at adversarial (range, h0)
#include "../cairo_b200/csrc/host/entropy.cpp"
#include <random>
#include <x86intrin.h>

static std::vector<uint8_t> plain_coder(const std::vector<uint64_t> &bins, size_t n, uint64_t &nbits)
{
    std::vector<uint8_t> out((n * 3) / 1 + 64, 0);      // generous: collapsed intervals emit 16 bits per bin
    uint64_t pos = 0;
    auto put = [&](uint32_t b) { if ((pos >> 3) >= out.size()) out.resize(out.size() * 2, 0); if (b) out[pos >> 3] |= (uint8_t) (1u << (pos & 7)); pos++; };
    uint32_t low = 0, high = 0xFFFF, e3 = 0, h0 = 1, tot = 2;
    auto emit = [&](uint32_t b) { put(b); for (; e3; --e3) put(b ^ 1u); };
    for (size_t i = 0; i < n; ++i)
    {
        const uint32_t bit = (uint32_t) ((bins[i >> 6] >> (i & 63)) & 1u);
        const uint32_t mid = low + (uint32_t) (((uint64_t) (high - low) * h0) / tot);
        if (bit) low = mid + 1; else { high = mid; h0++; }
        tot++;
        for (;;)
        {
            if (high <= 0x7FFF) emit(0);
            else if (low > 0x7FFF) { emit(1); low -= 0x8000; high -= 0x8000; }
            else if (low > 0x3FFF && high <= 3u * 0x3FFFu) { e3++; low -= 0x4000; high -= 0x4000; }
            else break;
            low = (low << 1) & 0xFFFF;
            high = ((high << 1) & 0xFFFF) | 1u;
        }
    }
    e3++;
    emit(low < 0x3FFF ? 0u : 1u);
    nbits = pos;
    out.resize((pos + 7) / 8);
    return out;
}

static int check()
{
    int bad = 0;
    std::mt19937_64 rng(7);
    // 1. the split-point identity: floor(r*h0/tot) == (r * (floor(h0*recip[tot] / 2^16) + 1)) >> 48
    {
        evx::reciprocal_table &rt = evx::recips();
        rt.grow(size_t(1) << 23);
        const uint64_t *rc = rt.data();
        uint64_t tested = 0;
        for (uint64_t tot = 2; tot < (uint64_t(1) << 23); ++tot)
        {
            // adversarial h0: extremes, multiples that make r*h0/tot an integer, and random ones
            uint64_t hs[6] = { 1, tot - 1, tot / 2, tot / 3 + 1, 1 + rng() % (tot - 1), 1 + rng() % (tot - 1) };
            uint64_t rs[6] = { 65535, 65534, 32768, 16384, rng() % 65536, 1 + rng() % 65535 };
            for (uint64_t h0 : hs) for (uint64_t r : rs)
            {
                const uint64_t F = (uint64_t) (((unsigned __int128) h0 * rc[tot]) >> 16) + 1;
                if (((r * F) >> 48) != (r * h0) / tot) { if (bad < 5) printf("identity fails: r %llu h0 %llu tot %llu\n", (unsigned long long) r, (unsigned long long) h0, (unsigned long long) tot); bad++; }
                tested++;
            }
            // exact multiples: r*h0 == m*tot
            for (int t = 0; t < 2 && tot < 65536; ++t)
            {
                const uint64_t r = tot, h0 = 1 + rng() % (tot - 1);
                const uint64_t F = (uint64_t) (((unsigned __int128) h0 * rc[tot]) >> 16) + 1;
                if (((r * F) >> 48) != (r * h0) / tot) bad++;
                tested++;
            }
        }
        printf("split-point identity: %llu cases, %d failures\n", (unsigned long long) tested, bad);
    }
    // 2. fast coder == plain coder
    struct { size_t n; double p0; } cases[] = { {0, .5}, {1, .5}, {63, .5}, {64, .3}, {65, .7}, {1000, .5}, {173000, .6}, {173000, .9}, {173000, .999},
                                               {173000, .001}, {400000, .5}, {3000000, .97}, {9500000, .6}, {9000000, .9999} };
    for (auto &c : cases)
    {
        std::vector<uint64_t> bins((c.n >> 6) + 2, 0);
        std::bernoulli_distribution d(1.0 - c.p0);
        for (size_t i = 0; i < c.n; ++i) if (d(rng)) bins[i >> 6] |= uint64_t(1) << (i & 63);
        // bursts: long runs flip the model and provoke long E3 chains
        if (c.n > 100000) for (size_t i = c.n / 3; i < c.n / 3 + 5000; ++i) bins[i >> 6] ^= uint64_t(1) << (i & 63);
        uint64_t nb = 0;
        std::vector<uint8_t> want = plain_coder(bins, c.n, nb), got;
        const uint64_t gb = evx::abac_encode_bins(bins.data(), c.n, got);
        const bool same = gb == nb && memcmp(got.data(), want.data(), want.size()) == 0;
        printf("n %zu p0 %.4f: %llu bits %s\n", c.n, c.p0, (unsigned long long) nb, same ? "ok" : "MISMATCH");
        if (!same) bad++;
    }
    return bad;
}

int main(int argc, char **argv)
{
    if (argc > 1 && !strcmp(argv[1], "check")) { int bad = check(); printf(bad ? "FAILED\n" : "all ok\n"); return bad ? 1 : 0; }
    size_t n = 173000;
    double p0 = argc > 1 ? atof(argv[1]) : 0.6;
    std::vector<uint64_t> bins((n >> 6) + 2, 0);
    std::mt19937 rng(1);
    std::bernoulli_distribution d(1.0 - p0);
    for (size_t i = 0; i < n; ++i) if (d(rng)) bins[i >> 6] |= uint64_t(1) << (i & 63);
    std::vector<uint8_t> out;
    uint64_t bits = 0, best = ~0ull; double bestns = 1e30;
    for (int rep = 0; rep < 30; ++rep)
    {
        auto t0 = std::chrono::steady_clock::now(); uint64_t c0 = __rdtsc();
        bits = evx::abac_encode_bins(bins.data(), n, out);
        uint64_t c1 = __rdtsc(); auto t1 = std::chrono::steady_clock::now();
        best = std::min(best, c1 - c0); bestns = std::min(bestns, std::chrono::duration<double, std::nano>(t1 - t0).count());
    }
    uint32_t h = 0; for (size_t i = 0; i < (bits + 7) / 8; ++i) h = h * 31 + out[i];
    printf("bins %zu bits %llu hash %08x  %.2f tsc/bin  %.2f ns/bin\n", n, (unsigned long long) bits, h, (double) best / n, bestns / n);
}
