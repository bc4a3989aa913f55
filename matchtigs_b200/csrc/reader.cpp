// reader.cpp -- host-side FASTA / bcalm2 record parser feeding the step API.
//
// Stands where genome-graph's io::fasta / io::bcalm2 readers stand in the reference
// (call sites src/bin.rs:896-899, 907-910): it only splits records, validates the bcalm ids and
// collects `L:<s>:<j>:<t>` links in file order; all graph construction happens on the device.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "matchtigs_b200.h"

struct mtg_unitigs {
    std::string seq;                 // concatenated sequences
    std::vector<uint64_t> offsets;   // [U+1]
    std::vector<uint64_t> link_a, link_b;
    std::vector<uint8_t> strand_a, strand_b;
};

namespace {
void set_err(char* errbuf, size_t errcap, const std::string& m) {
    if (errbuf && errcap) {
        size_t n = std::min(errcap - 1, m.size());
        memcpy(errbuf, m.data(), n);
        errbuf[n] = 0;
    }
}
}  // namespace

extern "C" {

int mtg_unitigs_parse(const char* text, size_t len, int bcalm, mtg_unitigs** out, char* errbuf, size_t errcap) {
    if (!out || (!text && len)) return MTG_ERR_INVALID;
    *out = nullptr;
    mtg_unitigs* u = new mtg_unitigs();
    u->seq.reserve(len);
    u->offsets.push_back(0);
    size_t i = 0;
    auto bad = [&](const std::string& m) {
        set_err(errbuf, errcap, m + " (record " + std::to_string(u->offsets.size() - 1) + ")");
        delete u;
        return (int)MTG_ERR_INPUT;
    };
    while (i < len) {
        while (i < len && (text[i] == '\n' || text[i] == '\r')) i++;
        if (i >= len) break;
        if (text[i] != '>') return bad("FASTA: expected '>'");
        const size_t hs = i + 1;
        const char* nl = (const char*)memchr(text + i, '\n', len - i);
        size_t he = nl ? (size_t)(nl - text) : len;
        i = he;
        if (he > hs && text[he - 1] == '\r') he--;
        const uint64_t rec = u->offsets.size() - 1;
        if (bcalm) {
            size_t p = hs;
            uint64_t id = 0;
            bool any = false;
            while (p < he && text[p] >= '0' && text[p] <= '9') {
                id = id * 10 + (uint64_t)(text[p] - '0');
                p++;
                any = true;
            }
            if (!any || id != rec) return bad("bcalm: record id != position");
            while (p < he) {
                while (p < he && (text[p] == ' ' || text[p] == '\t')) p++;
                size_t q = p;
                while (q < he && text[q] != ' ' && text[q] != '\t') q++;
                if (q - p >= 7 && text[p] == 'L' && text[p + 1] == ':') {  // L:<+/->:<id>:<+/->
                    const char s = text[p + 2];
                    size_t c = p + 4;
                    uint64_t j = 0;
                    bool digits = false;
                    while (c < q && text[c] >= '0' && text[c] <= '9') {
                        j = j * 10 + (uint64_t)(text[c] - '0');
                        c++;
                        digits = true;
                    }
                    if (text[p + 3] != ':' || !digits || c + 1 >= q || text[c] != ':') return bad("bcalm: malformed L field");
                    const char t = text[c + 1];
                    if ((s != '+' && s != '-') || (t != '+' && t != '-')) return bad("bcalm: malformed L sign");
                    u->link_a.push_back(rec);
                    u->strand_a.push_back(s == '+');
                    u->link_b.push_back(j);
                    u->strand_b.push_back(t == '+');
                }
                p = q;
            }
        }
        // sequence lines up to the next '>' at a line start
        while (i < len) {
            if (text[i] == '\n' || text[i] == '\r') {
                i++;
                continue;
            }
            if (text[i] == '>') break;
            const char* e = (const char*)memchr(text + i, '\n', len - i);
            size_t le = e ? (size_t)(e - text) : len;
            size_t ce = le;
            if (ce > i && text[ce - 1] == '\r') ce--;
            u->seq.append(text + i, ce - i);
            i = le;
        }
        u->offsets.push_back(u->seq.size());
    }
    *out = u;
    return MTG_OK;
}

void mtg_unitigs_free(mtg_unitigs* u) { delete u; }

int mtg_unitigs_view(const mtg_unitigs* u, const char** seq, const uint64_t** offsets, uint64_t* unitigs, const uint64_t** link_a,
                     const uint8_t** strand_a, const uint64_t** link_b, const uint8_t** strand_b, uint64_t* n_links) {
    if (!u) return MTG_ERR_INVALID;
    if (seq) *seq = u->seq.data();
    if (offsets) *offsets = u->offsets.data();
    if (unitigs) *unitigs = u->offsets.size() - 1;
    if (link_a) *link_a = u->link_a.data();
    if (strand_a) *strand_a = u->strand_a.data();
    if (link_b) *link_b = u->link_b.data();
    if (strand_b) *strand_b = u->strand_b.data();
    if (n_links) *n_links = u->link_a.size();
    return MTG_OK;
}

}  // extern "C"
