// reference_tests.cpp -- the reference's own tests for the hot path, restated against the C++
// host mirror (hostcpp/poly_b200.hpp) so that they read like the originals:
//   search/mash/mash_test.go:9-62, search/mash/example_test.go:9-22,
//   search/align/align_test.go:139-292, search/align/example_test.go:49-111,
//   primers/primers_test.go:29-84.
// Needs a B200: every call below lands in libpolyb200.so.  Run by tests/test_gpu_hostcpp.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "poly_b200.hpp"

static int failures = 0;
#define EXPECT(cond)                                                              \
    do {                                                                          \
        if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } \
    } while (0)

using namespace poly;

static void TestMash() {  // mash_test.go:9-62
    const std::string A = "ATGCGATCGATCGATCGATCGATCGATCGATCGATCGATCGATCGATCGATCGATCGATCGA";
    const std::string B = "ATCGATCGATCGATCGATCGATCGATCGATCGATCGAATGCGATCGATCGATCGATCGATCG";
    auto fingerprint1 = mash::New(17, 10); fingerprint1.Sketch(A);
    auto fingerprint2 = mash::New(17, 9); fingerprint2.Sketch(A);
    EXPECT(fingerprint1.Distance(fingerprint2) == 0);
    EXPECT(fingerprint2.Distance(fingerprint1) == 0);
    auto spoofed = mash::New(17, 10); spoofed.Sketches[0] = 0;
    EXPECT(fingerprint1.Distance(spoofed) == 1);
    spoofed = mash::New(17, 9);
    EXPECT(fingerprint1.Distance(spoofed) == 1);
    fingerprint1 = mash::New(17, 10); fingerprint1.Sketch(A);
    fingerprint2 = mash::New(17, 5); fingerprint2.Sketch(B);
    double d = fingerprint1.Distance(fingerprint2);
    EXPECT(d > 0.19 && d < 0.21 && d == 0.19999999999999996);
    fingerprint1 = mash::New(17, 10); fingerprint1.Sketch(B);
    fingerprint2 = mash::New(17, 5); fingerprint2.Sketch(A);
    EXPECT(fingerprint1.Distance(fingerprint2) == 0);
    // batched addition agrees with the object API
    auto batch = mash::SketchBatch({A, B, "ACGT", ""}, 17, 10);
    auto one = mash::New(17, 10); one.Sketch(B);
    EXPECT(batch[1].Sketches == one.Sketches);
    EXPECT(batch[2].Sketches == std::vector<uint32_t>(10, 0u));
    bool panicked = false;
    try { mash::New(17, 0).Distance(mash::New(17, 3)); } catch (const GoPanic &) { panicked = true; }
    EXPECT(panicked);
}

static void TestSmithWaterman() {  // align_test.go:139-292 (scores)
    auto alphabet = align::NewAlphabet({"-", "A", "C", "G", "T"});
    auto sub = align::NewSubstitutionMatrix(alphabet, alphabet,
                                            {{0, 0, 0, 0, 0}, {0, 3, -3, -3, -3}, {0, -3, 3, -3, -3}, {0, -3, -3, 3, -3}, {0, -3, -3, -3, 3}});
    auto scoring = align::NewScoring(&sub, -2);
    EXPECT(align::SmithWaterman("TGTTACGG", "GGTTGACTA", scoring) == 13);
    EXPECT(align::SmithWaterman("ACACACTA", "AGCACACA", scoring) == 17);
    EXPECT(align::SmithWaterman("", "GAT", scoring) == 0);
    EXPECT(align::SmithWaterman("", "", scoring) == 0);
    EXPECT(align::SmithWaterman("G", "A", scoring) == 0);
    EXPECT(align::SmithWaterman("G", "G", scoring) == 3);
    EXPECT(align::SmithWaterman("G", "GATTACA", scoring) == 3);
    // example_test.go:49-111
    auto a5 = align::NewAlphabet({"A", "C", "G", "T", "U"});
    std::vector<std::vector<int64_t>> m5(5, std::vector<int64_t>(5, -1));
    for (int i = 0; i < 5; ++i) m5[i][i] = 1;
    auto s5 = align::NewSubstitutionMatrix(a5, a5, m5);
    EXPECT(align::SmithWaterman("GATTACA", "GCATGCU", align::NewScoring(&s5, -1)) == 2);
    auto an = align::NewAlphabet({"A", "C", "G", "T", "-"});
    auto nuc4 = align::NewSubstitutionMatrix(an, an, {{0, 0, 0, 0, 0}, {0, 5, -4, -4, -4}, {0, -4, 5, -4, -4}, {0, -4, -4, 5, -4}, {0, -4, -4, -4, 5}});
    EXPECT(align::SmithWaterman("GATTACA", "GCATGCT", align::NewScoring(&nuc4, -1)) == 15);
    EXPECT(align::SmithWaterman("GATTACA", "GCATGCU", align::NewScoring(nullptr, -1)) == 2);
    // aligned strings (align_test.go:167-196, example_test.go:82,110)
    auto al = align::SmithWatermanAlign("TGTTACGG", "GGTTGACTA", scoring);
    EXPECT(al.score == 13 && al.alignA == "GTT-AC" && al.alignB == "GTTGAC");
    al = align::SmithWatermanAlign("ACACACTA", "AGCACACA", scoring);
    EXPECT(al.score == 17 && al.alignA == "A-CACACTA" && al.alignB == "AGCACAC-A");
    al = align::SmithWatermanAlign("GATTACA", "GCATGCT", align::NewScoring(&nuc4, -1));
    EXPECT(al.score == 15 && al.alignA == "GATTAC" && al.alignB == "GCATGC");
    al = align::SmithWatermanAlign("", "GAT", scoring);
    EXPECT(al.score == 0 && al.alignA.empty() && al.alignB.empty());
    // TestNeedlemanWunsch scores (align_test.go:11-137)
    auto nws = align::NewScoring(&s5, -1);
    EXPECT(align::NeedlemanWunschScore("GATTACA", "GCATGCU", nws) == 0);
    EXPECT(align::NeedlemanWunschScore("GATTACA", "GATTACA", nws) == 7);
    EXPECT(align::NeedlemanWunschScore("GATTACA", "GAT", nws) == -1);
    EXPECT(align::NeedlemanWunschScore("", "GAT", nws) == -3);
    EXPECT(align::NeedlemanWunschScore("G", "GATTACA", nws) == -5);
    al = align::NeedlemanWunschAlign("GATTACA", "GCATGCU", nws);  // example_test.go:10-47
    EXPECT(al.score == 0 && al.alignA == "G-ATTACA" && al.alignB == "GCA-TGCU");
    bool err = false;
    try { align::SmithWaterman("ACGT", "ACGX", scoring); } catch (const align::AlphabetError &e) { err = std::string(e.what()) == "Symbol X not in alphabet"; }
    EXPECT(err);
}

static void TestSantaLucia() {  // primers_test.go:29-84
    auto t = primers::SantaLucia("ACGATGGCAGTAGCATGC", 0.1e-6, 350e-3, 0.0);
    EXPECT(std::fabs(62.7 - t.meltingTemp) / 62.7 < 0.02);
    t = primers::SantaLucia("ACGTAGATCTACGT", 0.1e-6, 350e-3, 0.0);
    EXPECT(std::fabs(47.428514 - t.meltingTemp) / 47.428514 < 0.02);
    double tm = primers::MeltingTemp("GTAAAACGACGGCCAGT");
    EXPECT(std::fabs(52.8 - tm) / 52.8 < 0.02);
    EXPECT(std::fabs(tm - 52.63382276100299) <= 1e-6 * 52.63382276100299);
    bool panicked = false;
    try { primers::MeltingTemp(""); } catch (const GoPanic &) { panicked = true; }
    EXPECT(panicked);
}

static void TestFastaParser() {  // io/fasta/fasta_test.go:135-170,199-237
    using fasta::Fasta;
    for (bool alias : {true, false}) {
        auto r = fasta::ParseAll(">humen\nGATTACA\nCATGAT", 256, alias);  // EOF-ended Fasta not valid
        EXPECT(r.err_code == 0 && r.fastas.empty());
        r = fasta::ParseAll(">humen\nGATTACA\nCATGAT\n", 256, alias);
        EXPECT(r.err_code == 0 && r.fastas == std::vector<Fasta>({{"humen", "GATTACACATGAT"}}));
        r = fasta::ParseAll(">doggy or something\nGATTACA\n\nCATGAT\n>homunculus\nAAAA\n", 256, alias);
        EXPECT(r.err_code == 0 && r.fastas == std::vector<Fasta>({{"doggy or something", "GATTACACATGAT"}, {"homunculus", "AAAA"}}));
        r = fasta::Parse("testing\natagtagtagtagtagatgatgatgatgagatg\n\n\n\n\n\n\n\n\n\n\n", alias);  // TestReadEmptyFasta
        EXPECT(r.err_code != 0 && r.fastas.empty() && r.error() == "did not find fasta start '>', got to line 13");
        r = fasta::ParseAll(">0\n0123456789ABCDEF\n>1\nCAC\n", 2, alias);  // TestParseBufferFail
        EXPECT(r.err_code == 3);
        r = fasta::ParseAll(">OK Fasta\nABGABA\n>NotOKFasta\n", 2, alias);  // TestParseEOFAfterName
        EXPECT(r.err_code == 2 && r.fastas.size() == 1);
    }
}

int main() {
    check(pg_init(0));
    TestFastaParser();
    TestMash();
    TestSmithWaterman();
    TestSantaLucia();
    std::printf(failures ? "FAILED (%d)\n" : "ok: reference tests pass through the C++ host mirror\n", failures);
    return failures ? 1 : 0;
}
