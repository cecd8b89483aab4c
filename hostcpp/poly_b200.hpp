// poly_b200.hpp -- header-only C++ host mirror of the reference's Go API for the hot path,
// over the C ABI of libpolyb200.so (include/poly_b200.h).
//
// The reference is Go (compiled code) and no Go toolchain exists in this image, so the host
// side above the C ABI is written in C++ with the reference's names, argument meaning and
// failure behaviour:
//   poly::mash::New / Mash{KmerSize,SketchSize,Sketches} / Sketch / Similarity / Distance
//        <- /root/reference/search/mash/mash.go:52-140
//   poly::align::NewAlphabet / NewSubstitutionMatrix / NewScoring / SmithWaterman (score)
//        <- alphabet/alphabet.go:25-41, search/align/matrix/matrix.go:13-38,
//           search/align/align.go:73-95,171-203
//   poly::primers::SantaLucia / MeltingTemp <- primers/primers.go:70-105,121-128
//   poly::fasta::Parse / ParseAll            <- io/fasta/fasta.go:72-77,89-118,149-243
//   poly::pcr::DesignPrimers* / SimulateSimple / Simulate <- primers/pcr/pcr.go:44-186
//   poly::mash::ToJSON / FromJSON (the encoding/json shape of mash.Mash) and the PGSKETCH container
// A Go panic surfaces as poly::GoPanic; align's alphabet.Error as poly::align::AlphabetError.
// The Go package in go/ (cgo) is the real drop-in; this header is what the tests in this
// repository can compile and run (tests/test_gpu_hostcpp.py).  No compute happens here.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../include/poly_b200.h"

namespace poly {

struct GoPanic : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
    if (rc != PG_OK) throw Error(rc, std::string("libpolyb200: ") + pg_last_error());
}

// []string -> bytes + offsets (what the cgo shim does before crossing the ABI)
struct Flat {
    std::vector<uint8_t> bases;
    std::vector<uint64_t> offsets;
    explicit Flat(const std::vector<std::string> &seqs) {
        offsets.reserve(seqs.size() + 1);
        offsets.push_back(0);
        size_t total = 0;
        for (auto &s : seqs) total += s.size();
        bases.reserve(total);
        for (auto &s : seqs) {
            bases.insert(bases.end(), s.begin(), s.end());
            offsets.push_back(bases.size());
        }
    }
};

namespace mash {

// mash.Mash, mash.go:52-56
struct Mash {
    int KmerSize;
    int SketchSize;
    std::vector<uint32_t> Sketches;

    // (*Mash).Sketch, mash.go:68-104 (mutates the receiver)
    void Sketch(const std::string &sequence) {
        const int64_t n = (int64_t)sequence.size() - KmerSize;
        if (n <= 0) return;
        if (KmerSize < 0) throw GoPanic("slice bounds out of range");
        const uint64_t cnt = (uint64_t)std::min<int64_t>(n, SketchSize);
        std::vector<uint32_t> out(cnt ? cnt : 1);
        int32_t status = 0;
        int rc = pg_mash_sketch_uniform(reinterpret_cast<const uint8_t *>(sequence.data()), 1, (uint32_t)sequence.size(),
                                        KmerSize, SketchSize, 0, out.data(), cnt, &status);
        if (rc == PG_ERR_PANIC || status == PG_ITEM_PANIC) throw GoPanic("index out of range [-1]");  // mash.go:96-98
        check(rc);
        for (uint64_t i = 0; i < cnt; ++i) Sketches[i] = out[i];  // n < s: the tail is left untouched (mash.go:81-84)
    }

    struct Pair { int64_t same; double similarity, distance; };
    Pair pair(const Mash &other) const {
        if ((int)Sketches.size() < SketchSize || (int)other.Sketches.size() < other.SketchSize) throw GoPanic("index out of range");
        std::vector<uint32_t> sk(Sketches.begin(), Sketches.begin() + SketchSize);
        sk.insert(sk.end(), other.Sketches.begin(), other.Sketches.begin() + other.SketchSize);
        const uint64_t off[3] = {0, (uint64_t)SketchSize, (uint64_t)SketchSize + (uint64_t)other.SketchSize};
        const uint32_t a = 0, b = 1;
        Pair p{};
        int32_t st = 0;
        int rc = pg_mash_similarity_pairs(sk.data(), off, 2, &a, &b, 1, &p.same, &p.similarity, &p.distance, &st);
        if (rc == PG_ERR_PANIC) throw GoPanic("index out of range [-1]");
        check(rc);
        return p;
    }
    double Similarity(const Mash &other) const { return pair(other).similarity; }  // mash.go:107-135
    double Distance(const Mash &other) const { return pair(other).distance; }      // mash.go:138-140
};

// mash.New, mash.go:59-65
inline Mash New(int kmerSize, int sketchSize) {
    if (sketchSize < 0) throw GoPanic("makeslice: len out of range");
    return Mash{kmerSize, sketchSize, std::vector<uint32_t>((size_t)sketchSize, 0u)};
}

// batched addition: []string -> []Mash, each == New(k, s) + Sketch(seq).  One call, sharded over the
// GPUs of this process (devices empty: every visible GPU; reads are independent, mash.go:68-104).
inline std::vector<Mash> SketchBatch(const std::vector<std::string> &seqs, int k, int s, const std::vector<int32_t> &devices = {}) {
    Flat f(seqs);
    size_t maxlen = 0;
    for (auto &q : seqs) maxlen = std::max(maxlen, q.size());
    const uint64_t stride = std::max<int64_t>(1, std::min<int64_t>((int64_t)maxlen - k, s));
    std::vector<uint32_t> out(seqs.size() * stride), count(seqs.size());
    std::vector<int32_t> status(seqs.size());
    int rc = pg_mash_sketch_batch_multi(f.bases.data(), f.offsets.data(), seqs.size(), k, s, 0, out.data(), stride, count.data(), status.data(),
                                        devices.empty() ? nullptr : devices.data(), (int32_t)devices.size());
    if (rc != PG_OK && rc != PG_ERR_PANIC) check(rc);
    std::vector<Mash> res;
    res.reserve(seqs.size());
    for (size_t i = 0; i < seqs.size(); ++i) {
        if (status[i] == PG_ITEM_PANIC) throw GoPanic("index out of range [-1]");
        Mash m = New(k, s);
        for (uint32_t j = 0; j < count[i]; ++j) m.Sketches[j] = out[i * stride + j];
        res.push_back(std::move(m));
    }
    return res;
}

// In-place form: `slab` holds the Sketches arrays of seqs.size() Mash values back to back (n x s words).  Row i
// receives its min(len_i - k, s) words, the rest of the row is left as it is -- what (*Mash).Sketch does to an
// existing Mash (mash.go:73-80); a zeroed slab ends up as n fresh sketches.  Returns the per-read counts.
inline std::vector<uint32_t> SketchInto(const std::vector<std::string> &seqs, int k, int s, std::vector<uint32_t> &slab,
                                        const std::vector<int32_t> &devices = {}) {
    Flat f(seqs);
    if (slab.size() < seqs.size() * (size_t)std::max(s, 0)) throw std::invalid_argument("slab smaller than n x s");
    std::vector<uint32_t> count(seqs.size());
    std::vector<int32_t> status(seqs.size());
    uint32_t dummy = 0;
    int rc = pg_mash_sketch_batch_multi(f.bases.data(), f.offsets.data(), seqs.size(), k, s, PG_SKETCH_TAIL_KEEP,
                                        slab.empty() ? &dummy : slab.data(), (uint64_t)std::max(s, 0), count.data(), status.data(),
                                        devices.empty() ? nullptr : devices.data(), (int32_t)devices.size());
    if (rc != PG_OK && rc != PG_ERR_PANIC) check(rc);
    for (int32_t st : status)
        if (st == PG_ITEM_PANIC) throw GoPanic("index out of range [-1]");
    return count;
}

// batched addition: sketches of fixed-length reads + the all-pairs matrix on several GPUs of this process
// (sketch kernels store into every device's gathered buffer: fused all-gather; device r computes row block r)
struct SketchDistance {
    std::vector<uint32_t> sketches;  // n x s, full Go arrays
    std::vector<uint32_t> same;      // n x n matching counts, receiver = row
    std::vector<double> distance;    // n x n, 1 - same/s (mash.go:134,139)
};
inline SketchDistance SketchDistanceMulti(const std::string &reads, uint64_t n, uint32_t readLen, int k, int s,
                                          const std::vector<int32_t> &devices = {}) {
    SketchDistance r;
    r.sketches.resize(n * (uint64_t)s);
    r.same.resize(n * n);
    r.distance.resize(n * n);
    int rc = pg_mash_sketch_distance_multi(reinterpret_cast<const uint8_t *>(reads.data()), n, readLen, k, s,
                                           devices.empty() ? nullptr : devices.data(), (int32_t)devices.size(), r.sketches.data(), r.same.data(),
                                           r.distance.data());
    if (rc == PG_ERR_PANIC) throw GoPanic("index out of range [-1]");
    check(rc);
    return r;
}

// ---- sketch persistence -----------------------------------------------------------
// encoding/json of mash.Mash (mash.go:52-56): exported fields in declaration order, no spaces -- byte for
// byte what Go's json.Marshal emits, so single sketches interoperate with the reference.
inline std::string ToJSON(const Mash &m) {
    std::string out = "{\"KmerSize\":" + std::to_string(m.KmerSize) + ",\"SketchSize\":" + std::to_string(m.SketchSize) + ",\"Sketches\":";
    if (m.Sketches.empty() && m.SketchSize != 0) return out + "null}";  // nil slice
    out += "[";
    for (size_t i = 0; i < m.Sketches.size(); ++i) out += (i ? "," : "") + std::to_string(m.Sketches[i]);
    return out + "]}";
}
inline Mash FromJSON(const std::string &text) {
    auto number_after = [&](const char *key) -> long long {
        const size_t p = text.find(key);
        if (p == std::string::npos) throw std::invalid_argument(std::string("missing ") + key);
        return std::stoll(text.substr(p + std::strlen(key)));
    };
    Mash m{(int)number_after("\"KmerSize\":"), (int)number_after("\"SketchSize\":"), {}};
    size_t p = text.find("\"Sketches\":");
    if (p == std::string::npos) throw std::invalid_argument("missing Sketches");
    p += 11;
    if (text.compare(p, 4, "null") == 0) return m;
    if (text[p] != '[') throw std::invalid_argument("Sketches is not an array");
    ++p;
    while (p < text.size() && text[p] != ']') {
        size_t used = 0;
        m.Sketches.push_back((uint32_t)std::stoull(text.substr(p), &used));
        p += used;
        if (p < text.size() && text[p] == ',') ++p;
    }
    return m;
}

// PGSKETCH v1 container for sketch SETS (layout: poly_b200/sketchfile.py; little endian):
//   "PGSKETCH" | version 1 | KmerSize | SketchSize | flags (bit 0: dense) | n (u64) | words (u64) |
//   count[n] (absent when dense) | the informative words of sketch 0, 1, ... | CRC-32 of everything before
struct SketchSet {
    int KmerSize = 0, SketchSize = 0;
    std::vector<uint32_t> count;               // informative words per sketch
    std::vector<std::vector<uint32_t>> rows;   // rows[i].size() == count[i]
};
namespace detail {
inline uint32_t crc32(const uint8_t *p, size_t n, uint32_t crc = 0) {  // zlib's CRC-32 (reflected 0xEDB88320)
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return ~crc;
}
template <typename T>
inline void put(std::vector<uint8_t> &b, T v) {
    for (size_t i = 0; i < sizeof(T); ++i) b.push_back((uint8_t)(v >> (8 * i)));
}
template <typename T>
inline T get(const std::vector<uint8_t> &b, size_t &pos) {
    if (pos + sizeof(T) > b.size()) throw std::invalid_argument("truncated sketch file");
    T v = 0;
    for (size_t i = 0; i < sizeof(T); ++i) v |= (T)b[pos + i] << (8 * i);
    pos += sizeof(T);
    return v;
}
}  // namespace detail
inline std::vector<uint8_t> EncodeSketchSet(const SketchSet &set) {
    std::vector<uint8_t> b;
    const uint64_t n = set.rows.size();
    bool dense = n > 0;
    uint64_t words = 0;
    for (auto &r : set.rows) { dense = dense && (int)r.size() == set.SketchSize; words += r.size(); }
    b.insert(b.end(), {'P', 'G', 'S', 'K', 'E', 'T', 'C', 'H'});
    detail::put<uint32_t>(b, 1);
    detail::put<uint32_t>(b, (uint32_t)set.KmerSize);
    detail::put<uint32_t>(b, (uint32_t)set.SketchSize);
    detail::put<uint32_t>(b, dense ? 1u : 0u);
    detail::put<uint64_t>(b, n);
    detail::put<uint64_t>(b, words);
    if (!dense)
        for (auto &r : set.rows) detail::put<uint32_t>(b, (uint32_t)r.size());
    for (auto &r : set.rows)
        for (uint32_t w : r) detail::put<uint32_t>(b, w);
    detail::put<uint32_t>(b, detail::crc32(b.data(), b.size()));
    return b;
}
inline SketchSet DecodeSketchSet(const std::vector<uint8_t> &b) {
    if (b.size() < 44 || std::memcmp(b.data(), "PGSKETCH", 8) != 0) throw std::invalid_argument("not a PGSKETCH file");
    size_t tail = b.size() - 4, pos = 8;
    if (detail::crc32(b.data(), tail) != detail::get<uint32_t>(b, tail)) throw std::invalid_argument("sketch file checksum mismatch");
    if (detail::get<uint32_t>(b, pos) != 1) throw std::invalid_argument("not a PGSKETCH v1 file");
    SketchSet set;
    set.KmerSize = (int)detail::get<uint32_t>(b, pos);
    set.SketchSize = (int)detail::get<uint32_t>(b, pos);
    const uint32_t flags = detail::get<uint32_t>(b, pos);
    const uint64_t n = detail::get<uint64_t>(b, pos), words = detail::get<uint64_t>(b, pos);
    set.count.assign(n, (uint32_t)set.SketchSize);
    if (!(flags & 1u))
        for (uint64_t i = 0; i < n; ++i) set.count[i] = detail::get<uint32_t>(b, pos);
    uint64_t total = 0;
    for (uint32_t c : set.count) total += c;
    if (total != words || pos + 4 * words + 4 != b.size()) throw std::invalid_argument("sketch file is inconsistent");
    set.rows.resize(n);
    for (uint64_t i = 0; i < n; ++i) {
        set.rows[i].resize(set.count[i]);
        for (uint32_t j = 0; j < set.count[i]; ++j) set.rows[i][j] = detail::get<uint32_t>(b, pos);
    }
    return set;
}
inline void SaveSketchSet(const std::string &path, const SketchSet &set) {
    auto b = EncodeSketchSet(set);
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f || std::fwrite(b.data(), 1, b.size(), f) != b.size()) { if (f) std::fclose(f); throw std::runtime_error("cannot write " + path); }
    std::fclose(f);
}
inline SketchSet LoadSketchSet(const std::string &path) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot read " + path);
    std::vector<uint8_t> b;
    uint8_t buf[65536];
    for (size_t got; (got = std::fread(buf, 1, sizeof buf, f)) > 0;) b.insert(b.end(), buf, buf + got);
    std::fclose(f);
    return DecodeSketchSet(b);
}
// the set as Mash values (zero tail of a fresh Mash re-materialised)
inline std::vector<Mash> MashesOf(const SketchSet &set) {
    std::vector<Mash> out;
    for (auto &r : set.rows) {
        Mash m = New(set.KmerSize, set.SketchSize);
        std::copy(r.begin(), r.end(), m.Sketches.begin());
        out.push_back(std::move(m));
    }
    return out;
}

}  // namespace mash

namespace align {

struct AlphabetError : std::runtime_error {  // alphabet.Error, alphabet.go:14-22
    using std::runtime_error::runtime_error;
};

// Go's string(byte b): the UTF-8 encoding of code point b (align.go:90)
inline std::string go_string_of_byte(uint8_t b) {
    if (b < 0x80) return std::string(1, (char)b);
    return std::string{(char)(0xC0 | (b >> 6)), (char)(0x80 | (b & 0x3F))};
}

struct Alphabet {  // alphabet.go:9-12
    std::vector<std::string> symbols;
    std::map<std::string, int> encoding;
    int Encode(const std::string &symbol) const {  // alphabet.go:35-41
        auto it = encoding.find(symbol);
        if (it == encoding.end()) throw AlphabetError("Symbol " + symbol + " not in alphabet");
        return it->second;
    }
    std::vector<int16_t> byte_lut() const {
        std::vector<int16_t> lut(256, -1);
        for (int b = 0; b < 256; ++b) {
            auto it = encoding.find(go_string_of_byte((uint8_t)b));
            if (it != encoding.end()) lut[b] = (int16_t)it->second;
        }
        return lut;
    }
};
inline Alphabet NewAlphabet(const std::vector<std::string> &symbols) {  // alphabet.go:25-32
    Alphabet a;
    a.symbols = symbols;
    for (size_t i = 0; i < symbols.size(); ++i) a.encoding[symbols[i]] = (int)i;
    return a;
}

struct SubstitutionMatrix {  // matrix.go:13-17
    Alphabet FirstAlphabet, SecondAlphabet;
    std::vector<std::vector<int64_t>> scores;
    int64_t Score(const std::string &a, const std::string &b) const {  // matrix.go:28-38
        return scores[FirstAlphabet.Encode(a)][SecondAlphabet.Encode(b)];
    }
};
inline SubstitutionMatrix NewSubstitutionMatrix(const Alphabet &first, const Alphabet &second,
                                                const std::vector<std::vector<int64_t>> &scores) {
    if (first.symbols.size() != scores.size() || scores.empty() || second.symbols.size() != scores[0].size())
        throw std::invalid_argument("invalid dimensions of substitution matrix");  // matrix.go:21-23
    return SubstitutionMatrix{first, second, scores};
}
inline SubstitutionMatrix Default() {  // matrix.go:41-74: +1 / -1 over A..Z
    std::vector<std::string> letters;
    for (char c = 'A'; c <= 'Z'; ++c) letters.emplace_back(1, c);
    std::vector<std::vector<int64_t>> m(26, std::vector<int64_t>(26, -1));
    for (int i = 0; i < 26; ++i) m[i][i] = 1;
    return NewSubstitutionMatrix(NewAlphabet(letters), NewAlphabet(letters), m);
}

struct Scoring {  // align.go:73-76
    SubstitutionMatrix Matrix;
    int64_t GapPenalty;
};
inline Scoring NewScoring(const SubstitutionMatrix *m, int64_t gap) {  // align.go:79-87 (nil -> Default)
    return Scoring{m ? *m : Default(), gap};
}

// batched addition: score of SmithWaterman(query_i, template) for every query
inline std::vector<int64_t> SmithWatermanScores(const std::vector<std::string> &queries, const std::string &templ,
                                                const Scoring &sc, std::vector<std::string> *errors = nullptr) {
    Flat f(queries);
    auto lut_a = sc.Matrix.FirstAlphabet.byte_lut(), lut_b = sc.Matrix.SecondAlphabet.byte_lut();
    const int na = (int)sc.Matrix.scores.size(), nb = (int)sc.Matrix.scores[0].size();
    std::vector<int64_t> table((size_t)na * nb);
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) table[(size_t)i * nb + j] = sc.Matrix.scores[i][j];
    std::vector<int64_t> score(queries.size()), epos(queries.size());
    std::vector<int32_t> ecode(queries.size());
    check(pg_sw_score_batch(f.bases.data(), f.offsets.data(), queries.size(), reinterpret_cast<const uint8_t *>(templ.data()),
                            templ.size(), 1, lut_a.data(), lut_b.data(), table.data(), na, nb, sc.GapPenalty, score.data(),
                            ecode.data(), epos.data()));
    if (errors) {
        errors->assign(queries.size(), "");
        for (size_t i = 0; i < queries.size(); ++i)
            if (ecode[i]) {
                const uint8_t b = ecode[i] == 1 ? (uint8_t)queries[i][epos[i]] : (uint8_t)templ[epos[i]];
                (*errors)[i] = "Symbol " + go_string_of_byte(b) + " not in alphabet";  // alphabet.go:38
            }
    }
    return score;
}

// Score of align.SmithWaterman(stringA, stringB, scoring) (align.go:171-203).  The aligned
// strings (traceback, align.go:205-231) are a "next" row of SURVEY.md 8f.
inline int64_t SmithWaterman(const std::string &a, const std::string &b, const Scoring &sc) {
    std::vector<std::string> errs;
    auto s = SmithWatermanScores({a}, b, sc, &errs);
    if (!errs[0].empty()) throw AlphabetError(errs[0]);
    return s[0];
}

// Full align.SmithWaterman (align.go:171-232): score and the two aligned strings of the
// reference's traceback.  The shorter string (<= 64 symbols) rides in registers on the GPU.
struct Alignment {
    int64_t score;
    std::string alignA, alignB;
};
inline Alignment align_strings(bool global, const std::string &a, const std::string &b, const Scoring &sc) {
    const bool query_is_a = a.size() <= 64;
    const std::string &q = query_is_a ? a : b, &t = query_is_a ? b : a;
    Flat f({q});
    auto lut_a = sc.Matrix.FirstAlphabet.byte_lut(), lut_b = sc.Matrix.SecondAlphabet.byte_lut();
    const int na = (int)sc.Matrix.scores.size(), nb = (int)sc.Matrix.scores[0].size();
    std::vector<int64_t> table((size_t)na * nb);
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) table[(size_t)i * nb + j] = sc.Matrix.scores[i][j];
    uint64_t stride = global ? q.size() + t.size() + 8 : 2 * q.size() + 64;
    for (;;) {
        std::vector<uint8_t> oa(stride), ob(stride);
        int64_t score = 0, epos = 0;
        int32_t ecode = 0, status = 0;
        uint32_t len = 0;
        check((global ? pg_nw_align_batch : pg_sw_align_batch)(
            f.bases.data(), f.offsets.data(), 1, reinterpret_cast<const uint8_t *>(t.data()), t.size(), query_is_a ? 1 : 0,
            lut_a.data(), lut_b.data(), table.data(), na, nb, sc.GapPenalty, &score, &ecode, &epos, oa.data(), ob.data(), stride,
            &len, &status));
        if (ecode) {
            const uint8_t bad = ecode == 1 ? (uint8_t)a[epos] : (uint8_t)b[epos];
            throw AlphabetError("Symbol " + go_string_of_byte(bad) + " not in alphabet");
        }
        if (status == PG_ITEM_UNSUPPORTED) { stride = len + 8; continue; }
        return Alignment{score, std::string(oa.begin(), oa.begin() + len), std::string(ob.begin(), ob.begin() + len)};
    }
}

inline Alignment SmithWatermanAlign(const std::string &a, const std::string &b, const Scoring &sc) {
    return align_strings(false, a, b, sc);
}
// Full align.NeedlemanWunsch (align.go:100-166), including its loop condition (align.go:141).
inline Alignment NeedlemanWunschAlign(const std::string &a, const std::string &b, const Scoring &sc) {
    return align_strings(true, a, b, sc);
}

// Score of align.NeedlemanWunsch (align.go:100-134,166)
inline int64_t NeedlemanWunschScore(const std::string &a, const std::string &b, const Scoring &sc) {
    Flat f({a});
    auto lut_a = sc.Matrix.FirstAlphabet.byte_lut(), lut_b = sc.Matrix.SecondAlphabet.byte_lut();
    const int na = (int)sc.Matrix.scores.size(), nb = (int)sc.Matrix.scores[0].size();
    std::vector<int64_t> table((size_t)na * nb);
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) table[(size_t)i * nb + j] = sc.Matrix.scores[i][j];
    int64_t score = 0, epos = 0;
    int32_t ecode = 0;
    check(pg_nw_score_batch(f.bases.data(), f.offsets.data(), 1, reinterpret_cast<const uint8_t *>(b.data()), b.size(), 1,
                            lut_a.data(), lut_b.data(), table.data(), na, nb, sc.GapPenalty, &score, &ecode, &epos));
    if (ecode) {
        const uint8_t bad = ecode == 1 ? (uint8_t)a[epos] : (uint8_t)b[epos];
        throw AlphabetError("Symbol " + go_string_of_byte(bad) + " not in alphabet");
    }
    return score;
}

}  // namespace align

namespace primers {

struct Tm { double meltingTemp, dH, dS; };

// primers.SantaLucia, primers.go:70-105
inline Tm SantaLucia(const std::string &sequence, double primerConcentration, double saltConcentration,
                     double magnesiumConcentration) {
    const uint64_t off[2] = {0, sequence.size()};
    Tm t{};
    int32_t st = 0;
    int rc = pg_tm_batch(reinterpret_cast<const uint8_t *>(sequence.data()), off, 1, primerConcentration, saltConcentration,
                         magnesiumConcentration, &t.meltingTemp, &t.dH, &t.dS, &st);
    if (st == PG_ITEM_PANIC) throw GoPanic("index out of range [-1]");  // primers.go:89
    if (st == PG_ITEM_UNSUPPORTED) throw std::invalid_argument("byte >= 0x80 in primer (unsupported)");
    check(rc);
    return t;
}
// primers.MeltingTemp, primers.go:121-128
inline double MeltingTemp(const std::string &sequence) { return SantaLucia(sequence, 500e-9, 50e-3, 0.0).meltingTemp; }

inline std::vector<double> MeltingTemps(const std::vector<std::string> &seqs) {
    Flat f(seqs);
    std::vector<double> tm(seqs.size());
    std::vector<int32_t> st(seqs.size());
    int rc = pg_tm_batch(f.bases.data(), f.offsets.data(), seqs.size(), 500e-9, 50e-3, 0.0, tm.data(), nullptr, nullptr, st.data());
    for (auto s : st)
        if (s == PG_ITEM_PANIC) throw GoPanic("index out of range [-1]");
    check(rc);
    return tm;
}

}  // namespace primers

namespace transform {
// transform.ReverseComplement, transform/transform.go:15-23,78-109: bytes outside the table become 0
inline std::string ReverseComplement(const std::string &seq) {
    static const char *from = "ABCDGHKMNRSTVWYabcdghkmnrstvwy", *to = "TVGHCDMKNYSABWRtvghcdmknysabwr";
    uint8_t table[256] = {0};
    for (int i = 0; from[i]; ++i) table[(uint8_t)from[i]] = (uint8_t)to[i];
    std::string out(seq.size(), '\0');
    for (size_t i = 0; i < seq.size(); ++i) out[i] = (char)table[(uint8_t)seq[seq.size() - 1 - i]];
    return out;
}
}  // namespace transform

// primers/pcr (primers/pcr/pcr.go): the Tm searches (pcr.go:44-60, 93-100) and the binding-site search the
// reference does with a suffix array (pcr.go:87,110-115) run on the GPU; the fragment bookkeeping
// (pcr.go:117-166,181-195) is host work, as in the reference.
namespace pcr {

inline std::string upper_ascii(const std::string &s) {
    std::string u = s;
    for (char &c : u) {
        if ((uint8_t)c >= 0x80) throw std::invalid_argument("byte >= 0x80: strings.ToUpper on non-ASCII input is unsupported");
        if (c >= 'a' && c <= 'z') c = (char)(c - 32);
    }
    return u;
}

// pcr.DesignPrimersWithOverhangs, pcr.go:44-60
inline std::pair<std::string, std::string> DesignPrimersWithOverhangs(const std::string &sequence, const std::string &forwardOverhang,
                                                                       const std::string &reverseOverhang, double targetTm) {
    const std::string templ = upper_ascii(sequence);
    const uint64_t off[2] = {0, templ.size()};
    uint32_t fwd = 0, rev = 0;
    int32_t st = 0;
    int rc = pg_design_primers_batch(reinterpret_cast<const uint8_t *>(templ.data()), off, 1, targetTm, &fwd, &rev, &st);
    if (rc == PG_ERR_PANIC || st == PG_ITEM_PANIC) throw GoPanic("slice bounds out of range");  // the reference slices past the end
    check(rc);
    return {forwardOverhang + templ.substr(0, fwd), transform::ReverseComplement(reverseOverhang) + transform::ReverseComplement(templ.substr(templ.size() - rev))};
}
// pcr.DesignPrimers, pcr.go:62-66
inline std::pair<std::string, std::string> DesignPrimers(const std::string &sequence, double targetTm) {
    return DesignPrimersWithOverhangs(sequence, "", "", targetTm);
}

// pcr.SimulateSimple, pcr.go:73-169.  Like the reference it upper-cases primerList in place.
inline std::vector<std::string> SimulateSimple(const std::vector<std::string> &sequences, double targetTm, bool circular,
                                               std::vector<std::string> &primerList) {
    for (auto &p : primerList) p = upper_ascii(p);  // pcr.go:76-78
    std::vector<std::string> fragments;
    if (sequences.empty()) return fragments;
    std::vector<std::string> templates;
    for (auto &q : sequences) templates.push_back(upper_ascii(q));  // pcr.go:82
    // GPU pass 1: the minimal binding part of every primer (pcr.go:93-103)
    std::vector<std::string> minimal(primerList.size());
    std::vector<std::string> patterns;
    std::vector<std::pair<int, bool>> owner;  // (primer, is reverse complement)
    if (!primerList.empty()) {
        Flat f(primerList);
        std::vector<uint32_t> ml(primerList.size());
        std::vector<int32_t> st(primerList.size());
        int rc = pg_pcr_minimal_primer_batch(f.bases.data(), f.offsets.data(), primerList.size(), targetTm, ml.data(), st.data());
        for (int32_t x : st)
            if (x == PG_ITEM_PANIC) throw GoPanic("slice bounds out of range");  // primer shorter than 7 nt (pcr.go:35,96)
        if (rc != PG_ERR_PANIC) check(rc);
        for (size_t p = 0; p < primerList.size(); ++p) {
            const std::string part = primerList[p].substr(primerList[p].size() - ml[p]);
            if (part == primerList[p]) continue;  // pcr.go:103: ignored
            minimal[p] = part;
            patterns.push_back(part);
            owner.emplace_back((int)p, false);
            patterns.push_back(transform::ReverseComplement(part));
            owner.emplace_back((int)p, true);
        }
    }
    // GPU pass 2: every occurrence of every pattern in every template (pcr.go:110,113)
    struct Hit { uint32_t seq; uint64_t pos; uint32_t pat; };
    std::vector<Hit> hits;
    if (!patterns.empty()) {
        Flat ft(templates), fp(patterns);
        uint64_t cap = 1024, found = 0;
        for (;;) {
            std::vector<uint32_t> hs(cap), hq(cap);
            std::vector<uint64_t> hp(cap);
            int rc = pg_find_sites_batch(ft.bases.data(), ft.offsets.data(), templates.size(), fp.bases.data(), fp.offsets.data(),
                                         (uint32_t)patterns.size(), 0, hs.data(), hp.data(), hq.data(), cap, &found);
            if (rc == PG_ERR_ARG && found > cap) { cap = found; continue; }
            check(rc);
            for (uint64_t i = 0; i < found; ++i) hits.push_back({hs[i], hp[i], hq[i]});
            break;
        }
        // the order in which the reference fills its maps: by template, then primer (forward lookup first), then position
        std::sort(hits.begin(), hits.end(), [](const Hit &a, const Hit &b) {
            if (a.seq != b.seq) return a.seq < b.seq;
            if (a.pat != b.pat) return a.pat < b.pat;
            return a.pos < b.pos;
        });
    }
    size_t cursor = 0;
    for (size_t t = 0; t < templates.size(); ++t) {
        const std::string &templ = templates[t];
        std::map<uint64_t, std::vector<int>> fwd, rev;  // position -> primers bound there, in primer-list order; keys ascending
        for (; cursor < hits.size() && hits[cursor].seq == t; ++cursor)
            (owner[hits[cursor].pat].second ? rev : fwd)[hits[cursor].pos].push_back(owner[hits[cursor].pat].first);
        auto emit = [&](const std::string &text, uint64_t from, uint64_t to, const std::vector<int> &fps, const std::vector<int> &rps) {
            for (int fp : fps)      // generatePcrFragments, pcr.go:181-195
                for (int rp : rps)
                    fragments.push_back(primerList[fp].substr(0, primerList[fp].size() - minimal[fp].size()) + text.substr(from, to - from) +
                                        transform::ReverseComplement(primerList[rp]));
        };
        for (auto f = fwd.begin(); f != fwd.end(); ++f) {
            auto next = std::next(f);
            auto r = rev.upper_bound(f->first);  // reverse sites strictly to the right
            if (next != fwd.end()) {             // pcr.go:133-143: the first reverse site before the next forward site
                if (r != rev.end() && r->first < next->first) emit(templ, f->first, r->first, f->second, r->second);
                continue;
            }
            const bool found = r != rev.end();
            for (; r != rev.end(); ++r) emit(templ, f->first, r->first, f->second, r->second);  // pcr.go:146-151
            if (circular && !found) {            // pcr.go:153-164: look across the origin
                const std::string rotated = templ.substr(f->first) + templ.substr(0, f->first);
                for (auto q = rev.begin(); q != rev.end() && q->first < fwd.begin()->first; ++q)
                    emit(rotated, 0, templ.size() - f->first + q->first, f->second, q->second);
            }
        }
    }
    return fragments;
}

// pcr.Simulate, pcr.go:171-186: (fragments, error message; empty = nil)
struct Simulation {
    std::vector<std::string> fragments;
    bool hasFragments = true;  // false: the reference returns nil
    std::string error;
};
inline Simulation Simulate(const std::vector<std::string> &sequences, double targetTm, bool circular, std::vector<std::string> &primerList) {
    for (auto &p : primerList)
        if (p.size() < 7) return {{}, false, "Primers are too short."};  // minimalPrimerLength, pcr.go:35,174-178
    Simulation sim;
    sim.fragments = SimulateSimple(sequences, targetTm, circular, primerList);
    std::vector<std::string> again = primerList;
    again.insert(again.end(), sim.fragments.begin(), sim.fragments.end());
    if (SimulateSimple(sequences, targetTm, circular, again).size() != sim.fragments.size()) sim.error = "Concatemerization detected in PCR.";
    return sim;
}

}  // namespace pcr

namespace fasta {

// fasta.Fasta, io/fasta/fasta.go:66-69
struct Fasta {
    std::string Name, Sequence;
    bool operator==(const Fasta &o) const { return Name == o.Name && Sequence == o.Sequence; }
};
// the `error` half of ([]Fasta, error): code as pg_fasta_ingest reports it, 0 = nil
struct ParseResult {
    std::vector<Fasta> fastas;
    int32_t err_code = 0;
    uint64_t err_line = 0;
    std::string error() const {
        switch (err_code) {
            case 1: return "did not find fasta start '>', got to line " + std::to_string(err_line);
            case 2: return "empty fasta sequence, got to line " + std::to_string(err_line);
            case 3: return "line " + std::to_string(err_line) + " too large for buffer, use larger maxLineSize";
            case 4: return "bufio: buffer full";
            default: return "";
        }
    }
};

// NewParser(r, maxLineSize).ParseAll(), fasta.go:89-99; bufioAlias = give exactly what the reference
// gives over a strings.Reader / *os.File (see include/poly_b200.h, PG_FASTA_BUFIO_ALIAS)
inline ParseResult ParseAll(const std::string &text, uint32_t maxLineSize, bool bufioAlias = true) {
    uint64_t lines = 1;
    for (char c : text) lines += c == '\n';
    std::vector<uint8_t> bases(text.size() + 1), names(text.size() + 1);
    std::vector<uint64_t> off(lines + 1), noff(lines + 1);
    uint64_t n = 0, tot = 0, ntot = 0;
    ParseResult r;
    check(pg_fasta_ingest(reinterpret_cast<const uint8_t *>(text.data()), text.size(), maxLineSize,
                          bufioAlias ? PG_FASTA_BUFIO_ALIAS : 0u, bases.data(), bases.size(), off.data(), names.data(),
                          names.size(), noff.data(), lines, &n, &tot, &ntot, &r.err_code, &r.err_line));
    r.fastas.reserve(n);
    for (uint64_t i = 0; i < n; ++i)
        r.fastas.push_back({std::string(names.begin() + noff[i], names.begin() + noff[i + 1]),
                            std::string(bases.begin() + off[i], bases.begin() + off[i + 1])});
    return r;
}
// fasta.Parse, fasta.go:72-77
inline ParseResult Parse(const std::string &text, bool bufioAlias = true) { return ParseAll(text, 2 * 32 * 1024, bufioAlias); }

}  // namespace fasta

namespace fastq {

// fastq.Fastq, io/fastq/fastq.go:46-51
struct Fastq {
    std::string Identifier;
    std::map<std::string, std::string> Optionals;
    std::string Sequence, Quality;
};
// ([]Fastq, error) of fastq.Parse: the records before the first one the reference rejects, and that error
// (code as pg_fastq_ingest reports it, 0 = nil; err_line = the 1-based line the parser stopped at)
struct ParseResult {
    std::vector<Fastq> fastqs;
    int32_t err_code = 0;
    uint64_t err_line = 0;
    std::string error() const {
        const std::string l = std::to_string(err_line);
        switch (err_code) {
            case 1: return "line " + l + " failed: unexepcted EOF encountered";  // sic, fastq.go:150
            case 2: return "empty fastq sequence, got to line " + l;
            case 3: return "empty quality sequence, got to line " + l;
            case 4: return "did not find fastq start '@', got to line " + l;
            case 5: return "reference panics (index out of range) while parsing the identifier at line " + l;
            case 6: return "line " + l + " too large for buffer, use larger maxLineSize";
            default: return "";
        }
    }
};

// fastq.Parse, io/fastq/fastq.go:54-59 (ParseNext :117-214): records on the GPU (line index, per-record checks in the
// reference's order, dense sequences); Identifier / Optionals / Quality are cut out of `text` with the line spans.
inline ParseResult Parse(const std::string &text) {
    uint64_t cap = 1;
    for (char c : text) cap += c == '\n';
    cap = cap / 4 + 1;
    std::vector<uint8_t> bases(text.size() + 1);
    std::vector<uint64_t> off(cap + 1), spans(4 * (cap + 1));
    uint64_t n = 0, tot = 0;
    ParseResult r;
    check(pg_fastq_ingest_records(reinterpret_cast<const uint8_t *>(text.data()), text.size(), bases.data(), bases.size(), off.data(),
                                  spans.data(), cap, &n, &tot, &r.err_code, &r.err_line));
    r.fastqs.reserve(n);
    for (uint64_t i = 0; i < n; ++i) {
        Fastq f;
        const std::string line = text.substr(spans[4 * i], spans[4 * i + 1]);
        size_t pos = 0;
        bool first = true;
        while (pos <= line.size()) {  // strings.Split(line, " "), fastq.go:157
            size_t sp = line.find(' ', pos);
            if (sp == std::string::npos) sp = line.size();
            const std::string tok = line.substr(pos, sp - pos);
            if (first) {
                f.Identifier = tok.substr(1);  // without the '@', fastq.go:158
                first = false;
            } else {  // "key=value": strings.Split(datum, "=")[0], [1] (every datum holds '=': checked on the GPU)
                const size_t eq = tok.find('=');
                const size_t eq2 = tok.find('=', eq + 1);
                f.Optionals[tok.substr(0, eq)] = tok.substr(eq + 1, eq2 == std::string::npos ? std::string::npos : eq2 - eq - 1);
            }
            pos = sp + 1;
        }
        f.Sequence.assign(bases.begin() + off[i], bases.begin() + off[i + 1]);
        f.Quality = text.substr(spans[4 * i + 2], spans[4 * i + 3]);
        r.fastqs.push_back(std::move(f));
    }
    return r;
}

}  // namespace fastq
}  // namespace poly
