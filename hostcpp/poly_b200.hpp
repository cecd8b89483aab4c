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
// A Go panic surfaces as poly::GoPanic; align's alphabet.Error as poly::align::AlphabetError.
// The Go package in go/ (cgo) is the real drop-in; this header is what the tests in this
// repository can compile and run (tests/test_gpu_hostcpp.py).  No compute happens here.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
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

// batched addition: []string -> []Mash, each == New(k, s) + Sketch(seq)
inline std::vector<Mash> SketchBatch(const std::vector<std::string> &seqs, int k, int s) {
    Flat f(seqs);
    size_t maxlen = 0;
    for (auto &q : seqs) maxlen = std::max(maxlen, q.size());
    const uint64_t stride = std::max<int64_t>(1, std::min<int64_t>((int64_t)maxlen - k, s));
    std::vector<uint32_t> out(seqs.size() * stride), count(seqs.size());
    std::vector<int32_t> status(seqs.size());
    int rc = pg_mash_sketch_batch(f.bases.data(), f.offsets.data(), seqs.size(), k, s, 0, out.data(), stride, count.data(), status.data());
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
}  // namespace poly
