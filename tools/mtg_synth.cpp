// mtg_synth.cpp -- TEST/BENCH TOOLING (not product, not oracle): deterministic synthetic genomes
// and an exact compacted-de-Bruijn-graph unitig builder that emits bcalm2-style FASTA
// (`>id LN:i:len L:+:j:- ...`), because BCALM2/GGCAT are not available in this image and every
// BASELINE.json config starts from unitigs.
//
// k must be odd and <= 63 (k <= 31: 64-bit k-mers, else 128-bit).  Output is deterministic for a
// given input regardless of thread count: unitigs are ordered by their smallest canonical end k-mer.
#include <omp.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {
using u8 = uint8_t;
using u32 = uint32_t;
using u64 = uint64_t;
using u128 = unsigned __int128;

struct Rng {  // splitmix64: identical streams on every platform
    u64 s;
    explicit Rng(u64 seed) : s(seed) {}
    u64 next() {
        u64 z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    u64 below(u64 n) { return n ? next() % n : 0; }
    double unit() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};
const char ACGT[5] = "ACGT";
inline int code(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

// ---------------- genome generators ----------------
void random_seq(std::string& s, size_t len, Rng& r) {
    s.resize(len);
    size_t i = 0;
    while (i < len) {
        u64 x = r.next();
        for (int j = 0; j < 32 && i < len; j++, x >>= 2) s[i++] = ACGT[x & 3];
    }
}
char mutate(char c, Rng& r) { return ACGT[(code(c) + 1 + r.below(3)) & 3]; }

// Uniform background with repeat families pasted over it (configs 1-3, 5).
std::string make_genome(size_t len, u64 seed, u32 families, u32 copies, u32 min_len, u32 max_len, double divergence,
                        u32 tandem_arrays) {
    Rng r(seed);
    std::string g;
    random_seq(g, len, r);
    for (u32 f = 0; f < families; f++) {
        size_t L = min_len + r.below(max_len - min_len + 1);
        if (L + 1 >= len) continue;
        std::string fam;
        random_seq(fam, L, r);
        for (u32 c = 0; c < copies; c++) {
            size_t pos = r.below(len - L);
            bool rc = r.next() & 1;
            for (size_t i = 0; i < L; i++) {
                char ch = rc ? ACGT[3 - code(fam[L - 1 - i])] : fam[i];
                if (divergence > 0 && r.unit() < divergence) ch = mutate(ch, r);
                g[pos + i] = ch;
            }
        }
    }
    for (u32 t = 0; t < tandem_arrays; t++) {
        size_t unit = 2 + r.below(60), reps = 5 + r.below(40);
        if (unit * reps + 1 >= len) continue;
        size_t pos = r.below(len - unit * reps);
        for (size_t i = unit; i < unit * reps; i++) g[pos + i] = g[pos + i % unit];
    }
    return g;
}

// Pangenome: strains derived from one ancestor through a shared pool of variant sites, each
// with a population frequency, so that strains share variants (bubbles in the union graph).
std::vector<std::string> make_pangenome(const std::string& anc, u32 strains, u64 seed, double snp_site_rate,
                                        double indel_site_rate, double private_snp_rate) {
    Rng r(seed);
    struct Site {
        size_t pos;
        u8 kind;  // 0 snp, 1 insertion, 2 deletion
        char alt;
        u32 len;
        float freq;
        std::string ins;
    };
    std::vector<Site> sites;
    size_t n = anc.size();
    for (size_t p = 0; p < n; p++) {
        double u = r.unit();
        if (u < snp_site_rate) {
            Site s{p, 0, mutate(anc[p], r), 1, 0, {}};
            double f = r.unit();
            s.freq = (float)(0.02 + 0.96 * f * f);  // skewed towards rare
            sites.push_back(s);
        } else if (u < snp_site_rate + indel_site_rate) {
            Site s{p, (u8)(1 + (r.next() & 1)), 'A', 1, 0, {}};
            u32 L = 1;
            while (r.unit() < 0.7 && L < 2000) L++;
            if (r.unit() < 0.02) L = 200 + (u32)r.below(3000);  // a few kb-scale events
            s.len = L;
            if (s.kind == 1) random_seq(s.ins, L, r);
            double f = r.unit();
            s.freq = (float)(0.02 + 0.96 * f * f);
            sites.push_back(s);
        }
    }
    std::vector<std::string> out(strains);
    for (u32 st = 0; st < strains; st++) {
        Rng rs(seed * 1000003ull + 100 + st);
        std::string& g = out[st];
        g.reserve(n + n / 50);
        size_t si = 0, p = 0;
        while (p < n) {
            while (si < sites.size() && sites[si].pos < p) si++;
            if (si < sites.size() && sites[si].pos == p && rs.unit() < sites[si].freq) {
                const Site& s = sites[si];
                if (s.kind == 0) {
                    g.push_back(s.alt);
                    p++;
                } else if (s.kind == 1) {
                    g.push_back(anc[p]);
                    g += s.ins;
                    p++;
                } else {
                    p += s.len;
                }
                continue;
            }
            char c = anc[p++];
            if (private_snp_rate > 0 && rs.unit() < private_snp_rate) c = mutate(c, rs);
            g.push_back(c);
        }
    }
    return out;
}

// ---------------- k-mer helpers ----------------
template <class K>
struct KmerOps {
    int k;
    K mask;
    explicit KmerOps(int k_) : k(k_) { mask = (k * 2 == (int)sizeof(K) * 8) ? ~(K)0 : (((K)1 << (2 * k)) - 1); }
    K rc(K x) const {
        K r = 0;
        for (int i = 0; i < k; i++) {
            r = (r << 2) | (3 - (x & 3));
            x >>= 2;
        }
        return r;
    }
    K canon(K x, bool* flipped = nullptr) const {
        K r = rc(x);
        if (flipped) *flipped = r < x;
        return r < x ? r : x;
    }
    K succ(K x, int c) const { return ((x << 2) | (K)c) & mask; }
};

template <class K>
struct Dbg {
    KmerOps<K> ops;
    std::vector<K> kmers;       // sorted distinct canonical k-mers
    std::vector<u64> bucket;    // prefix index over the top bits
    int bshift = 0, bbits = 0;
    std::vector<u8> adj;        // bits 0-3: successors of the canonical orientation, 4-7: of its reverse complement
    explicit Dbg(int k) : ops(k) {}

    void build_index() {
        size_t n = kmers.size();
        bbits = 1;
        while (((size_t)1 << bbits) < n / 4 + 1 && bbits < 26) bbits++;
        bbits = std::min(bbits, 2 * ops.k);
        bshift = 2 * ops.k - bbits;
        bucket.assign(((size_t)1 << bbits) + 1, 0);
        for (size_t i = 0; i < n; i++) bucket[(size_t)(kmers[i] >> bshift) + 1]++;
        for (size_t b = 0; b + 1 < bucket.size(); b++) bucket[b + 1] += bucket[b];
    }
    // index of canonical k-mer c, or -1
    int64_t find(K c) const {
        size_t b = (size_t)(c >> bshift);
        size_t lo = bucket[b], hi = bucket[b + 1];
        while (lo < hi) {
            size_t mid = (lo + hi) >> 1;
            if (kmers[mid] < c) lo = mid + 1;
            else hi = mid;
        }
        return (lo < bucket[b + 1] && kmers[lo] == c) ? (int64_t)lo : -1;
    }
    struct Ori {  // oriented k-mer: canonical index + whether the oriented form is the reverse complement of the canonical
        int64_t idx;
        bool flip;
    };
    K oriented(Ori y) const { return y.flip ? ops.rc(kmers[y.idx]) : kmers[y.idx]; }
    int outdeg(Ori y) const { return __builtin_popcount((adj[y.idx] >> (y.flip ? 4 : 0)) & 15); }
    int indeg(Ori y) const { return outdeg(Ori{y.idx, !y.flip}); }
    Ori step(Ori y, int c) const {
        bool fl;
        K z = ops.canon(ops.succ(oriented(y), c), &fl);
        return Ori{find(z), fl};
    }
    Ori only_succ(Ori y) const {
        int bits = (adj[y.idx] >> (y.flip ? 4 : 0)) & 15;
        return step(y, __builtin_ctz(bits));
    }
    // may the unitig containing y (as its last k-mer so far) be extended forward?
    bool extends(Ori y, Ori* next) const {
        if (outdeg(y) != 1) return false;
        Ori z = only_succ(y);
        if (z.idx == y.idx) return false;  // hairpin / self loop: a k-mer may appear once per unitig
        if (indeg(z) != 1) return false;
        *next = z;
        return true;
    }
};

struct Unitig {
    std::string seq;
    u64 first_idx, last_idx;  // canonical indices of the first/last k-mer
    bool first_flip, last_flip;
};

template <class K>
std::string build_unitigs_t(const std::vector<std::string>& seqs, int k, int threads, u64* n_kmers_out, u64* n_unitigs_out) {
    Dbg<K> d(k);
    const KmerOps<K>& ops = d.ops;
    if (threads > 0) omp_set_num_threads(threads);
    // 1. all canonical k-mers
    std::vector<size_t> off(seqs.size() + 1, 0);
    for (size_t s = 0; s < seqs.size(); s++) off[s + 1] = off[s] + (seqs[s].size() >= (size_t)k ? seqs[s].size() - k + 1 : 0);
    std::vector<K> all(off.back());
    for (size_t s = 0; s < seqs.size(); s++) {
        const std::string& q = seqs[s];
        if (q.size() < (size_t)k) continue;
        size_t nk = q.size() - k + 1;
        const size_t CH = 1 << 20;
#pragma omp parallel for schedule(dynamic)
        for (size_t c0 = 0; c0 < nk; c0 += CH) {
            size_t c1 = std::min(nk, c0 + CH);
            K f = 0, r = 0;
            for (int i = 0; i < k - 1; i++) {
                int c = code(q[c0 + i]);
                f = (f << 2) | (K)c;
                r = (r >> 2) | ((K)(3 - c) << (2 * (k - 1)));
            }
            for (size_t p = c0; p < c1; p++) {
                int c = code(q[p + k - 1]);
                f = ((f << 2) | (K)c) & ops.mask;
                r = (r >> 2) | ((K)(3 - c) << (2 * (k - 1)));
                all[off[s] + p] = f < r ? f : r;
            }
        }
    }
    // 2. sort + unique (bucket by top byte, sort buckets in parallel)
    {
        const int TB = 12;
        int sh = std::max(0, 2 * k - TB);
        size_t nb = (size_t)1 << std::min(TB, 2 * k);
        std::vector<size_t> cnt(nb + 1, 0);
        for (size_t i = 0; i < all.size(); i++) cnt[(size_t)(all[i] >> sh) + 1]++;
        for (size_t b = 0; b < nb; b++) cnt[b + 1] += cnt[b];
        std::vector<K> tmp(all.size());
        {
            std::vector<size_t> pos(cnt.begin(), cnt.end() - 1);
            for (size_t i = 0; i < all.size(); i++) tmp[pos[(size_t)(all[i] >> sh)]++] = all[i];
        }
        all.clear();
        all.shrink_to_fit();
        std::vector<size_t> ucnt(nb, 0);
#pragma omp parallel for schedule(dynamic, 8)
        for (size_t b = 0; b < nb; b++) {
            std::sort(tmp.begin() + cnt[b], tmp.begin() + cnt[b + 1]);
            ucnt[b] = std::unique(tmp.begin() + cnt[b], tmp.begin() + cnt[b + 1]) - (tmp.begin() + cnt[b]);
        }
        size_t total = 0;
        for (size_t b = 0; b < nb; b++) total += ucnt[b];
        d.kmers.resize(total);
        size_t w = 0;
        for (size_t b = 0; b < nb; b++) {
            std::copy(tmp.begin() + cnt[b], tmp.begin() + cnt[b] + ucnt[b], d.kmers.begin() + w);
            w += ucnt[b];
        }
    }
    const size_t G = d.kmers.size();
    *n_kmers_out = G;
    d.build_index();
    // 3. adjacency bits
    d.adj.assign(G, 0);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < G; i++) {
        u8 a = 0;
        K x = d.kmers[i], xr = ops.rc(x);
        for (int c = 0; c < 4; c++) {
            if (d.find(ops.canon(ops.succ(x, c))) >= 0) a |= (u8)(1 << c);
            if (d.find(ops.canon(ops.succ(xr, c))) >= 0) a |= (u8)(16 << c);
        }
        d.adj[i] = a;
    }
    // 4. unitigs: walk forward from every oriented start; keep the copy whose start end is the
    //    smaller (canonical index, flip) of its two ends.
    using Ori = typename Dbg<K>::Ori;
    auto is_start = [&](Ori y) -> bool {  // nothing may extend into y from behind
        Ori back{y.idx, !y.flip};          // walking backwards == walking forward on the reverse complement
        Ori nx;
        return !d.extends(back, &nx);
    };
    std::vector<u8> visited(G, 0);
    std::vector<std::vector<Unitig>> per_thread(omp_get_max_threads());
    auto walk = [&](Ori start, Unitig& u) {
        K x = d.oriented(start);
        u.seq.resize(k);
        for (int i = 0; i < k; i++) u.seq[i] = ACGT[(int)((x >> (2 * (k - 1 - i))) & 3)];
        Ori cur = start, nx;
        visited[cur.idx] = 1;
        while (d.extends(cur, &nx)) {
            if (nx.idx == start.idx) break;  // closed an isolated cycle
            K z = d.oriented(nx);
            u.seq.push_back(ACGT[(int)(z & 3)]);
            cur = nx;
            visited[cur.idx] = 1;
        }
        u.first_idx = start.idx;
        u.first_flip = start.flip;
        u.last_idx = cur.idx;
        u.last_flip = cur.flip;
    };
#pragma omp parallel for schedule(dynamic, 4096)
    for (size_t i = 0; i < G; i++) {
        for (int fl = 0; fl < 2; fl++) {
            Ori y{(int64_t)i, (bool)fl};
            if (!is_start(y)) continue;
            // find the other end cheaply first to decide ownership without building the string twice
            Ori cur = y, nx;
            while (d.extends(cur, &nx)) cur = nx;
            Ori other{cur.idx, !cur.flip};  // start of the reverse-complement copy
            bool keep = (y.idx < other.idx) || (y.idx == other.idx && (int)y.flip <= (int)other.flip);
            if (!keep) continue;
            Unitig u;
            walk(y, u);
            per_thread[omp_get_thread_num()].push_back(std::move(u));
        }
    }
    std::vector<Unitig> unitigs;
    for (auto& v : per_thread) {
        for (auto& u : v) unitigs.push_back(std::move(u));
        v.clear();
    }
    // isolated cycles: every k-mer on them has in = out = 1 and no start
    for (size_t i = 0; i < G; i++) {
        if (visited[i]) continue;
        Unitig u;
        walk(Ori{(int64_t)i, false}, u);
        unitigs.push_back(std::move(u));
    }
    std::sort(unitigs.begin(), unitigs.end(), [](const Unitig& a, const Unitig& b) {
        if (a.first_idx != b.first_idx) return a.first_idx < b.first_idx;
        return a.first_flip < b.first_flip;
    });
    *n_unitigs_out = unitigs.size();
    // 5. links.  end_of[i] = unitig whose first (bit0) / last (bit1) k-mer is canonical k-mer i.
    std::vector<u32> owner(G, 0xFFFFFFFFu);
    for (size_t u = 0; u < unitigs.size(); u++) {
        owner[unitigs[u].first_idx] = (u32)u;
        owner[unitigs[u].last_idx] = (u32)u;
    }
    std::vector<std::string> headers(unitigs.size());
#pragma omp parallel for schedule(dynamic, 1024)
    for (size_t u = 0; u < unitigs.size(); u++) {
        const Unitig& t = unitigs[u];
        std::string h = ">" + std::to_string(u) + " LN:i:" + std::to_string(t.seq.size());
        for (int side = 0; side < 2; side++) {
            // side 0: leave through the last k-mer in + orientation; side 1: leave through rc(first k-mer), i.e. "-"
            Ori y = side == 0 ? Ori{(int64_t)t.last_idx, t.last_flip} : Ori{(int64_t)t.first_idx, !t.first_flip};
            int bits = (d.adj[y.idx] >> (y.flip ? 4 : 0)) & 15;
            for (int c = 0; c < 4; c++) {
                if (!(bits >> c & 1)) continue;
                Ori z = d.step(y, c);
                u32 v = owner[z.idx];
                if (v == 0xFFFFFFFFu) continue;  // cannot happen for a correct compaction
                const Unitig& tv = unitigs[v];
                char sign;
                if (tv.first_idx == (u64)z.idx && tv.first_flip == z.flip) sign = '+';
                else if (tv.last_idx == (u64)z.idx && tv.last_flip == !z.flip) sign = '-';
                else continue;
                h += std::string(" L:") + (side == 0 ? '+' : '-') + ":" + std::to_string(v) + ":" + sign;
            }
        }
        headers[u] = std::move(h);
    }
    std::string out;
    size_t total = 0;
    for (size_t u = 0; u < unitigs.size(); u++) total += headers[u].size() + unitigs[u].seq.size() + 2;
    out.reserve(total);
    for (size_t u = 0; u < unitigs.size(); u++) {
        out += headers[u];
        out += '\n';
        out += unitigs[u].seq;
        out += '\n';
    }
    return out;
}

std::vector<std::string> split_lines(const char* text, size_t len) {
    std::vector<std::string> v;
    size_t i = 0;
    while (i < len) {
        size_t j = i;
        while (j < len && text[j] != '\n') j++;
        if (j > i && text[i] != '>') v.emplace_back(text + i, text + j);
        i = j + 1;
    }
    return v;
}

char* dup_out(const std::string& s, size_t* len) {
    char* p = (char*)std::malloc(s.size() + 1);
    std::memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    *len = s.size();
    return p;
}
}  // namespace

extern "C" {
void mts_free(char* p) { std::free(p); }

// One random genome with repeat families; returns a malloc'd ACGT string.
char* mts_genome(size_t len, uint64_t seed, uint32_t families, uint32_t copies, uint32_t min_len, uint32_t max_len,
                 double divergence, uint32_t tandem_arrays, size_t* out_len) {
    return dup_out(make_genome(len, seed, families, copies, min_len, max_len, divergence, tandem_arrays), out_len);
}

// Strains of one ancestor, newline separated.
char* mts_pangenome(const char* ancestor, size_t len, uint32_t strains, uint64_t seed, double snp_site_rate,
                    double indel_site_rate, double private_snp_rate, size_t* out_len) {
    std::string anc(ancestor, len);
    auto v = make_pangenome(anc, strains, seed, snp_site_rate, indel_site_rate, private_snp_rate);
    std::string out;
    for (auto& s : v) {
        out += s;
        out += '\n';
    }
    return dup_out(out, out_len);
}

// Input: sequences separated by newlines (lines starting with '>' are ignored).  Output: bcalm2-style FASTA.
char* mts_unitigs(const char* text, size_t len, int k, int threads, size_t* out_len, uint64_t* n_kmers, uint64_t* n_unitigs) {
    if (k < 3 || k > 63 || k % 2 == 0) return nullptr;
    auto seqs = split_lines(text, len);
    for (auto& s : seqs)
        for (char c : s)
            if (code(c) < 0) return nullptr;
    std::string out = k <= 31 ? build_unitigs_t<u64>(seqs, k, threads, n_kmers, n_unitigs)
                              : build_unitigs_t<u128>(seqs, k, threads, n_kmers, n_unitigs);
    return dup_out(out, out_len);
}
}
