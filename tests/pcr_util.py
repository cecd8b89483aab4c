"""CPU restatement of pcr.SimulateSimple / Simulate (primers/pcr/pcr.go:73-195) for the tests: Tm from
the oracle (po_melting_temp), binding sites by bytes.find (what suffixarray.Lookup(p, -1) returns:
every occurrence, overlapping ones included), fragment assembly as the reference orders it."""
import oracle_ffi as o

_PAIRS = dict(zip(b"ABCDGHKMNRSTVWYabcdghkmnrstvwy", b"TVGHCDMKNYSABWRtvghcdmknysabwr"))


def revcomp(s: bytes) -> bytes:
    """transform.ReverseComplement: bytes outside the table become 0 (transform/transform.go:15-23,78-109)."""
    return bytes(_PAIRS.get(c, 0) for c in reversed(s))


def occurrences(text: bytes, pat: bytes):
    if not pat:
        return []
    out, i = [], text.find(pat)
    while i >= 0:
        out.append(i)
        i = text.find(pat, i + 1)
    return out


def minimal_length(primer: bytes, target: float) -> int:
    """pcr.go:93-100; raises IndexError where Go panics."""
    if len(primer) < 7:            # minimalPrimerLength, pcr.go:35
        raise IndexError("slice bounds out of range")
    minimal, index = 0, 7
    while o.melting_temp(primer[len(primer) - index:]) < target:
        minimal = index
        if index == len(primer):
            break
        index += 1
    return minimal


def simulate_simple(sequences, target, circular, primer_list):
    primers = [bytes(p).upper() for p in primer_list]
    frags = []
    for seq in sequences:
        seq = bytes(seq).upper()
        fwd, rev, minimal = {}, {}, [b""] * len(primers)
        for pi, primer in enumerate(primers):
            ml = minimal_length(primer, target)
            mp = primer[len(primer) - ml:]
            if mp == primer:
                continue
            minimal[pi] = mp
            for loc in occurrences(seq, mp):
                fwd.setdefault(loc, []).append(pi)
            for loc in occurrences(seq, revcomp(mp)):
                rev.setdefault(loc, []).append(pi)

        def gen(s, f, r, fis, ris):
            return [primers[fi][: len(primers[fi]) - len(minimal[fi])] + s[f:r] + revcomp(primers[ri]) for fi in fis for ri in ris]

        fl, rl = sorted(fwd), sorted(rev)
        for idx, f in enumerate(fl):
            if idx + 1 != len(fl):
                hit = next((r for r in rl if f < r < fl[idx + 1]), None)
                if hit is not None:
                    frags += gen(seq, f, hit, fwd[f], rev[hit])
                continue
            later = [r for r in rl if f < r]
            for r in later:
                frags += gen(seq, f, r, fwd[f], rev[r])
            if circular and not later:
                for r in rl:
                    if fl[0] > r:
                        frags += gen(seq[f:] + seq[:f], 0, len(seq) - f + r, fwd[f], rev[r])
    return [f.decode("latin-1") for f in frags]


def simulate(sequences, target, circular, primer_list):
    if any(len(p) < 7 for p in primer_list):      # pcr.go:174-178
        return None, "Primers are too short."
    first = simulate_simple(sequences, target, circular, primer_list)
    second = simulate_simple(sequences, target, circular, list(primer_list) + [f.encode() for f in first])
    return first, ("Concatemerization detected in PCR." if len(first) != len(second) else None)


# ---- the reference's own expectations (primers/pcr/pcr_test.go, example_test.go) ----------------
GENE = (b"aataattacaccgagataacacatcatggataaaccgatactcaaagattctatgaagctatttgaggcacttggtacgatcaagtcgcgctcaatgtttggtggcttcggacttttcgc"
        b"tgatgaaacgatgtttgcactggttgtgaatgatcaacttcacatacgagcagaccagcaaacttcatctaacttcgagaagcaagggctaaaaccgtacgtttataaaaagcgtggttttcc"
        b"agtcgttactaagtactacgcgatttccgacgacttgtgggaatccagtgaacgcttgatagaagtagcgaagaagtcgttagaacaagccaatttggaaaaaaagcaacaggcaagtagtaa"
        b"gcccgacaggttgaaagacctgcctaacttacgactagcgactgaacgaatgcttaagaaagctggtataaaatcagttgaacaacttgaagagaaaggtgcattgaatgcttacaaagcgat"
        b"acgtgactctcactccgcaaaagtaagtattgagctactctgggctttagaaggagcgataaacggcacgcactggagcgtcgttcctcaatctcgcagagaagagctggaaaatgcgctttc"
        b"ttaa")
BAD_FRAGMENT = (b"ATGACCATGATTACGCCAAGCTTGCATGCCTGCAGGTCGACTCTAGAGGATCCCCGGGTACCGAGCTCGAATTCACTGGCCGTCGTTTTACAACGTCGTGACTGGGAAAACCCTGGCG"
                b"TTACCCAACTTAATCGCCTTGCAGCACATCCCCCTTTCGCCAGCTGGCGTAATAGCGAAGAGGCCCGCACCGATCGCCCTTCCCAACAGTTGCGCAGCCTGAATGGCGAATGGCGCCTGA"
                b"TGCGGTATTTTCTCCTTACGCATCTGTGCGGTATTTCACACCGCATATGGTGCACTCTCAGTACAATCTGCTCTGATGCCGCATAG")
FWD, REV = b"TTATAGGTCTCATACTAATAATTACACCGAGATAACACATCATGG", b"TATATGGTCTCTTCATTTAAGAAAGCGCATTTTCCAGC"
# example_test.go:68 (ExampleSimulate) == pcr_test.go:96 (TestIssue279PCRBug `want`)
FULL_AMPLICON = (FWD + GENE[29:].upper() + b"ATGAAGAGACCATATA").decode()
CIRCULAR_TARGET = ("ACTCTGGGCTTTAGAAGGAGCGATAAACGGCACGCACTGGAGCGTCGTTCCTCAATCTCGCAGAGAAGAGCTGGAAAATGCGCTTTCTTAAAATAATTACACCGAGATAACACATCATG"
                   "GATAAACCGATACTCAAAGATTCTATGAAGCTATTTGAGGCACTT")   # pcr_test.go:59


def reference_cases():
    """(sequences, circular, primers, check(fragments, error))"""
    yield [GENE], False, [FWD, REV], lambda fr, e: fr == [FULL_AMPLICON] and e is None                     # ExampleSimulate
    yield [GENE, BAD_FRAGMENT], False, [FWD, REV], lambda fr, e: len(fr) == 1                               # Example_basic
    yield [GENE], False, [REV, FWD, b"CTGCAGGTCGACTCTAG"], lambda fr, e: fr == [FULL_AMPLICON] and e is None  # TestSimulatePrimerRejection, TestIssue279PCRBug
    yield ([GENE], False, [b"gatactcaaagattctatgaagctatttgaggcacttggtacg", b"tatcgctttgtaagcattcaatgcacctttctcttcaagttg",
                           b"gtcgttcctcaatctcgcagagaagagctggaaaatg"], lambda fr, e: len(fr) == 1)           # TestSimulateMoreThanOneForward
    yield ([GENE], True, [b"actctgggctttagaaggagcgataaacggc", b"aagtgcctcaaatagcttcatagaatctttgagtatcgg"],
           lambda fr, e: fr[0] == CIRCULAR_TARGET)                                                          # TestSimulateCircular
    yield ([GENE], False, [b"AATAATTACACCGAGATAACACATCATGG", b"CCATGATGTGTTATCTCGGTGTAATTATTTTAAGAAAGCGCATTTTCCAGC"],
           lambda fr, e: e is not None)                                                                     # TestSimulateConcatemerization
    yield [GENE], False, [FWD, b"ACGT"], lambda fr, e: fr is None and e == "Primers are too short."        # pcr.go:174-178
