// trajectory.cpp -- see trajectory.hpp.  XTC layout: SURVEY.md Appendix A.
#include "trajectory.hpp"

#include <atomic>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <thread>

namespace mdsctk_cli {

namespace {

constexpr int kMagic = 1995;
constexpr int kFirstIdx = 9;
const int kMagicInts[] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406,
    512, 645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642,
    26007, 32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127,
    524287, 660561, 832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491,
    6658042, 8388607, 10568983, 13316085, 16777216};
constexpr int kLastIdx = (int)(sizeof(kMagicInts) / sizeof(kMagicInts[0]));

inline int32_t be32(const unsigned char *p) { return (int32_t)((uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]); }
inline float bef(const unsigned char *p) { uint32_t u = (uint32_t)be32(p); float f; std::memcpy(&f, &u, 4); return f; }

// MSB-first bit reader with a 64-bit window.
struct Bits {
    const unsigned char *p, *end;
    uint64_t window = 0;
    int have = 0;
    uint32_t take(int n)  // n <= 32
    {
        while (have < n) { window = (window << 8) | (p < end ? *p : 0); ++p; have += 8; }
        have -= n;
        return (uint32_t)((window >> have) & ((n == 32) ? 0xffffffffull : ((1ull << n) - 1)));
    }
    // value = sum byte_j * 256^j over successive 8-bit reads (the last one short), as three mixed-radix digits
    void triple(int nbits, const unsigned sizes[3], int out[3])
    {
        unsigned __int128 v = 0;
        int shift = 0;
        for (; nbits > 8; nbits -= 8, shift += 8) v |= (unsigned __int128)take(8) << shift;
        if (nbits > 0) v |= (unsigned __int128)take(nbits) << shift;
        out[2] = (int)(v % sizes[2]); v /= sizes[2];
        out[1] = (int)(v % sizes[1]);
        out[0] = (int)(v / sizes[1]);
    }
};

int bitlen(unsigned __int128 v) { int n = 0; for (; v; v >>= 1) ++n; return n; }

// Size in bytes of the frame starting at p (header walk), or 0 if malformed / truncated.
size_t frame_size(const unsigned char *p, size_t left, int *natoms)
{
    if (left < 56 || be32(p) != kMagic) return 0;
    const int n = be32(p + 4);
    if (n <= 0 || be32(p + 52) != n) return 0;
    *natoms = n;
    if (n <= 9) return 56 + (size_t)n * 12 <= left ? 56 + (size_t)n * 12 : 0;
    if (left < 92) return 0;
    const int nbytes = be32(p + 88);
    if (nbytes < 0) return 0;
    size_t total = 92 + (((size_t)nbytes + 3) & ~(size_t)3);
    if (92 + (size_t)nbytes > left) return 0;
    return total > left ? left : total;  // a final frame may lack its padding
}

bool decode_frame(const unsigned char *p, int natoms, float *out)
{
    if (natoms <= 9) {
        for (int i = 0; i < natoms * 3; ++i) out[i] = bef(p + 56 + 4 * i);
        return true;
    }
    const float inv = 1.0f / bef(p + 56);
    int lo[3], hi[3];
    for (int d = 0; d < 3; ++d) { lo[d] = be32(p + 60 + 4 * d); hi[d] = be32(p + 72 + 4 * d); }
    int smallidx = be32(p + 84);
    const int nbytes = be32(p + 88);
    if (smallidx < kFirstIdx || smallidx >= kLastIdx) return false;
    unsigned span[3], bits_each[3] = {0, 0, 0};
    for (int d = 0; d < 3; ++d) span[d] = (unsigned)(hi[d] - lo[d] + 1);
    int bits_all;
    if ((span[0] | span[1] | span[2]) > 0xffffffu) {
        for (int d = 0; d < 3; ++d) bits_each[d] = (unsigned)bitlen(span[d]);
        bits_all = 0;
    } else {
        bits_all = bitlen((unsigned __int128)span[0] * span[1] * span[2]);
    }
    int smaller = kMagicInts[smallidx - 1 > kFirstIdx ? smallidx - 1 : kFirstIdx] / 2;
    int smallnum = kMagicInts[smallidx] / 2;
    unsigned small_span[3] = {(unsigned)kMagicInts[smallidx], (unsigned)kMagicInts[smallidx], (unsigned)kMagicInts[smallidx]};
    Bits br{p + 92, p + 92 + nbytes};
    int run = 0, written = 0, cur[3], prev[3];
    auto emit = [&](const int v[3]) {
        if (written >= natoms) return false;
        for (int d = 0; d < 3; ++d) out[3 * written + d] = (float)v[d] * inv;
        ++written;
        return true;
    };
    for (int i = 0; i < natoms;) {
        if (bits_all == 0) for (int d = 0; d < 3; ++d) cur[d] = (int)br.take((int)bits_each[d]);
        else br.triple(bits_all, span, cur);
        ++i;
        for (int d = 0; d < 3; ++d) { cur[d] += lo[d]; prev[d] = cur[d]; }
        int change = 0;
        if (br.take(1)) {
            run = (int)br.take(5);
            change = run % 3;
            run -= change;
            --change;
        }
        if (run > 0) {
            for (int k = 0; k < run; k += 3) {
                br.triple(smallidx, small_span, cur);
                ++i;
                for (int d = 0; d < 3; ++d) cur[d] += prev[d] - smallnum;
                if (k == 0) {  // the group's second atom is stored first
                    for (int d = 0; d < 3; ++d) std::swap(cur[d], prev[d]);
                    if (!emit(prev)) return false;
                } else {
                    for (int d = 0; d < 3; ++d) prev[d] = cur[d];
                }
                if (!emit(cur)) return false;
            }
        } else if (!emit(cur)) {
            return false;
        }
        smallidx += change;
        if (smallidx < kFirstIdx || smallidx >= kLastIdx) return false;
        if (change < 0) { smallnum = smaller; smaller = smallidx > kFirstIdx ? kMagicInts[smallidx - 1] / 2 : 0; }
        else if (change > 0) { smaller = smallnum; smallnum = kMagicInts[smallidx] / 2; }
        small_span[0] = small_span[1] = small_span[2] = (unsigned)kMagicInts[smallidx];
    }
    return written == natoms;
}

struct MassEntry { const char *name; float mass; };
// element symbols in upper case (PDB / GRO atom names are upper case); values of GROMACS' atommass.dat
const MassEntry kTwoLetter[] = {
    {"CL", 35.45300f}, {"BR", 79.90000f}, {"NA", 22.98970f}, {"MG", 24.30500f}, {"CA", 40.08000f},
    {"FE", 55.84700f}, {"ZN", 65.37000f}, {"CU", 63.54600f}, {"SI", 28.08000f}, {"AL", 26.98150f}, {"MN", 54.93800f}};
const MassEntry kOneLetter[] = {
    {"H", 1.00790f},   {"C", 12.01070f},  {"N", 14.00670f},  {"O", 15.99940f},  {"S", 32.06500f},
    {"P", 30.97380f},  {"F", 18.99840f},  {"B", 10.81100f},  {"I", 126.90450f}, {"K", 39.10200f}};

std::string upper_trim(const std::string &in)
{
    std::string out;
    for (char ch : in)
        if (!std::isspace((unsigned char)ch)) out.push_back((char)std::toupper((unsigned char)ch));
    return out;
}

bool lookup(const MassEntry *tab, size_t n, const std::string &sym, float *m)
{
    for (size_t i = 0; i < n; ++i)
        if (sym == tab[i].name) { *m = tab[i].mass; return true; }
    return false;
}

// Mass of an atom from what a PDB / GRO file says about it (the reference gets it from GROMACS' atommass lookup,
// knn_rms.cpp:150-153).  `element`: the PDB element columns 77-78 when present; `two_letter`: the name is known to
// start with a two-letter element symbol (PDB: the name starts in column 13; GRO: residue name == atom name, i.e. an
// ion) -- that is what tells calcium "CA  " from the alpha carbon " CA ".  Unknown names get carbon's mass, with a
// warning, instead of silently changing the centring.
float mass_of(const std::string &raw_name, const std::string &element, bool two_letter)
{
    float m = 0.f;
    const std::string el = upper_trim(element);
    if (!el.empty() && (lookup(kTwoLetter, sizeof kTwoLetter / sizeof *kTwoLetter, el, &m) ||
                        lookup(kOneLetter, sizeof kOneLetter / sizeof *kOneLetter, el, &m)))
        return m;
    std::string name = upper_trim(raw_name);
    size_t b = 0;
    while (b < name.size() && std::isdigit((unsigned char)name[b])) ++b;     // "1HD1" style hydrogens
    name = name.substr(b);
    if (two_letter && name.size() >= 2 && lookup(kTwoLetter, sizeof kTwoLetter / sizeof *kTwoLetter, name.substr(0, 2), &m)) return m;
    if (!name.empty() && lookup(kOneLetter, sizeof kOneLetter / sizeof *kOneLetter, name.substr(0, 1), &m)) return m;
    static int warned = 0;
    if (warned++ < 8) std::cerr << "WARNING: no mass known for atom name '" << raw_name << "': using 12.011 (give --mass-file)" << std::endl;
    return 12.011f;
}

}  // namespace

bool XtcFile::open(const std::string &path, std::string *err)
{
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) { *err = "cannot open " + path; return false; }
    const std::streamsize n = f.tellg();
    f.seekg(0);
    bytes.resize((size_t)(n > 0 ? n : 0));
    if (n > 0 && !f.read(reinterpret_cast<char *>(bytes.data()), n)) { *err = "cannot read " + path; return false; }
    frame_offset.clear();
    natoms = 0;
    size_t off = 0;
    while (off < bytes.size()) {
        int na = 0;
        const size_t sz = frame_size(bytes.data() + off, bytes.size() - off, &na);
        if (sz == 0 || (natoms && na != natoms)) break;   // like read_next_xtc: stop at the first bad frame
        natoms = na;
        frame_offset.push_back(off);
        off += sz;
    }
    if (frame_offset.empty()) { *err = path + " holds no readable xtc frame"; return false; }
    return true;
}

bool XtcFile::decode_all(float *xyz, int nthreads, std::string *err) const
{
    const long long n = frames();
    if (is_flat) { std::memcpy(xyz, bytes.data(), (size_t)n * natoms * 12); return true; }
    if (nthreads < 1) nthreads = 1;
    std::atomic<long long> next(0);
    std::atomic<long long> bad(-1);
    auto work = [&]() {
        for (;;) {
            const long long f0 = next.fetch_add(256);
            if (f0 >= n) return;
            for (long long f = f0; f < std::min(n, f0 + 256); ++f)
                if (!decode_frame(bytes.data() + frame_offset[(size_t)f], natoms, xyz + (size_t)f * natoms * 3))
                    bad.store(f);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    if (bad.load() >= 0) { *err = "corrupt xtc frame " + std::to_string(bad.load()); return false; }
    return true;
}

bool read_topology_masses(const std::string &path, std::vector<float> *mass, std::string *err)
{
    std::ifstream f(path);
    if (!f) { *err = "cannot open " + path; return false; }
    mass->clear();
    std::string line;
    const bool gro = path.size() >= 4 && path.compare(path.size() - 4, 4, ".gro") == 0;
    if (path.size() >= 4 && path.compare(path.size() - 4, 4, ".tpr") == 0) {
        *err = ".tpr topologies need GROMACS; give a .pdb/.gro (masses from atom names) or --mass-file";
        return false;
    }
    if (gro) {
        std::getline(f, line);
        std::getline(f, line);
        const int n = std::atoi(line.c_str());
        for (int i = 0; i < n && std::getline(f, line); ++i) {
            if (line.size() < 15) break;
            // residue name == atom name: a monoatomic ion ("NA", "CL", "ZN", "CA" the calcium ion, ...)
            mass->push_back(mass_of(line.substr(10, 5), "", upper_trim(line.substr(5, 5)) == upper_trim(line.substr(10, 5))));
        }
    } else {
        while (std::getline(f, line)) {
            if (line.compare(0, 6, "ENDMDL") == 0) break;   // first model only
            if (line.compare(0, 4, "ATOM") != 0 && line.compare(0, 6, "HETATM") != 0) continue;
            if (line.size() < 16) continue;
            // a name that starts in column 13 is a two-letter element unless it is a four-character hydrogen name
            const bool col13 = std::isalpha((unsigned char)line[12]) && !(line[12] == 'H' && line[15] != ' ');
            mass->push_back(mass_of(line.substr(12, 4), line.size() >= 78 ? line.substr(76, 2) : std::string(), col13));
        }
    }
    if (mass->empty()) { *err = "no atoms found in " + path; return false; }
    return true;
}

bool read_mass_file(const std::string &path, std::vector<float> *mass, std::string *err)
{
    std::ifstream f(path);
    if (!f) { *err = "cannot open " + path; return false; }
    mass->clear();
    float m;
    while (f >> m) mass->push_back(m);
    if (mass->empty()) { *err = "no masses in " + path; return false; }
    return true;
}

}  // namespace mdsctk_cli
