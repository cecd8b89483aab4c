// Package align: drop-in for github.com/bebop/poly/search/align.SmithWaterman
// (search/align/align.go:171-232: score, aligned strings, error) and for the score of
// NeedlemanWunsch (align.go:100-134,166), backed by libpolyb200.so.  NOT COMPILED HERE.
// Both need the shorter string to have <= 64 symbols; longer pairs keep the pure-Go path.
package align

import (
	"fmt"

	"github.com/bebop/poly/alphabet"
	"github.com/bebop/poly/internal/polyb200"
	"github.com/bebop/poly/search/align/matrix"
)

// Scoring mirrors align.Scoring (align.go:73-76).
type Scoring struct {
	SubstitutionMatrix *matrix.SubstitutionMatrix
	GapPenalty         int
}

// NewScoring mirrors align.NewScoring (align.go:79-87).
func NewScoring(substitutionMatrix *matrix.SubstitutionMatrix, gapPenalty int) (Scoring, error) {
	if substitutionMatrix == nil {
		substitutionMatrix = matrix.Default
	}
	return Scoring{SubstitutionMatrix: substitutionMatrix, GapPenalty: gapPenalty}, nil
}

// flatten turns the interface-keyed alphabets into byte LUTs + a dense table (host side, once per call).
func flatten(m *matrix.SubstitutionMatrix) (lutA, lutB [256]int16, table []int64, nA, nB int) {
	symA, symB := m.FirstAlphabet.Symbols(), m.SecondAlphabet.Symbols()
	nA, nB = len(symA), len(symB)
	for b := 0; b < 256; b++ {
		lutA[b], lutB[b] = -1, -1
		if c, err := m.FirstAlphabet.Encode(string(byte(b))); err == nil {
			lutA[b] = int16(c)
		}
		if c, err := m.SecondAlphabet.Encode(string(byte(b))); err == nil {
			lutB[b] = int16(c)
		}
	}
	table = make([]int64, nA*nB)
	for i := 0; i < nA; i++ {
		for j := 0; j < nB; j++ {
			v, _ := m.Score(symA[i], symB[j])
			table[i*nB+j] = int64(v)
		}
	}
	return
}

// SmithWatermanScores scores SmithWaterman(query, template, scoring) for every query in one GPU pass.
func SmithWatermanScores(queries []string, template string, scoring Scoring) ([]int, []error) {
	lutA, lutB, table, nA, nB := flatten(scoring.SubstitutionMatrix)
	bases, offsets := polyb200.Flatten(queries)
	score, ec, ep, err := polyb200.SWScoreBatch(bases, offsets, template, true, &lutA, &lutB, table, nA, nB, int64(scoring.GapPenalty))
	if err != nil {
		panic(err)
	}
	scores, errs := make([]int, len(queries)), make([]error, len(queries))
	for i := range queries {
		scores[i] = int(score[i])
		if ec[i] == 1 {
			_, errs[i] = scoring.SubstitutionMatrix.FirstAlphabet.Encode(string(queries[i][ep[i]]))
		} else if ec[i] == 2 {
			_, errs[i] = scoring.SubstitutionMatrix.SecondAlphabet.Encode(string(template[ep[i]]))
		}
	}
	return scores, errs
}

// SmithWatermanScore is the score (and error) of align.SmithWaterman(stringA, stringB, scoring).
func SmithWatermanScore(stringA, stringB string, scoring Scoring) (int, error) {
	s, e := SmithWatermanScores([]string{stringA}, stringB, scoring)
	if e[0] != nil {
		return 0, e[0]
	}
	return s[0], nil
}

// SmithWaterman keeps the reference signature (align.go:171): score, aligned strings, error.
// The shorter of the two strings must have <= 64 symbols (it rides in registers on the GPU);
// longer pairs should keep using the pure-Go implementation.
func SmithWaterman(stringA string, stringB string, scoring Scoring) (int, string, string, error) {
	lutA, lutB, table, nA, nB := flatten(scoring.SubstitutionMatrix)
	queryIsA := len(stringA) <= 64
	q, t := stringA, stringB
	if !queryIsA {
		q, t = stringB, stringA
	}
	bases, offsets := polyb200.Flatten([]string{q})
	score, ec, ep, alignA, alignB, err := polyb200.SWAlignBatch(bases, offsets, t, queryIsA, &lutA, &lutB, table, nA, nB,
		int64(scoring.GapPenalty), 2*len(q)+64)
	if err != nil {
		panic(err)
	}
	if ec[0] == 1 {
		_, e := scoring.SubstitutionMatrix.FirstAlphabet.Encode(string(stringA[ep[0]]))
		return 0, "", "", e
	} else if ec[0] == 2 {
		_, e := scoring.SubstitutionMatrix.SecondAlphabet.Encode(string(stringB[ep[0]]))
		return 0, "", "", e
	}
	return int(score[0]), alignA[0], alignB[0], nil
}

// NeedlemanWunsch keeps the reference signature (align.go:100): score, aligned strings, error.
func NeedlemanWunsch(stringA string, stringB string, scoring Scoring) (int, string, string, error) {
	lutA, lutB, table, nA, nB := flatten(scoring.SubstitutionMatrix)
	queryIsA := len(stringA) <= 64
	q, t := stringA, stringB
	if !queryIsA {
		q, t = stringB, stringA
	}
	bases, offsets := polyb200.Flatten([]string{q})
	score, ec, ep, alignA, alignB, err := polyb200.NWAlignBatch(bases, offsets, t, queryIsA, &lutA, &lutB, table, nA, nB,
		int64(scoring.GapPenalty), len(q)+len(t)+8)
	if err != nil {
		panic(err)
	}
	if ec[0] == 1 {
		_, e := scoring.SubstitutionMatrix.FirstAlphabet.Encode(string(stringA[ep[0]]))
		return 0, "", "", e
	} else if ec[0] == 2 {
		_, e := scoring.SubstitutionMatrix.SecondAlphabet.Encode(string(stringB[ep[0]]))
		return 0, "", "", e
	}
	return int(score[0]), alignA[0], alignB[0], nil
}

// NeedlemanWunschScore is the score of align.NeedlemanWunsch(stringA, stringB, scoring).
func NeedlemanWunschScore(stringA, stringB string, scoring Scoring) (int, error) {
	lutA, lutB, table, nA, nB := flatten(scoring.SubstitutionMatrix)
	bases, offsets := polyb200.Flatten([]string{stringA})
	score, ec, ep, err := polyb200.NWScoreBatch(bases, offsets, stringB, true, &lutA, &lutB, table, nA, nB, int64(scoring.GapPenalty))
	if err != nil {
		panic(err)
	}
	if ec[0] == 1 {
		_, e := scoring.SubstitutionMatrix.FirstAlphabet.Encode(string(stringA[ep[0]]))
		return 0, e
	} else if ec[0] == 2 {
		_, e := scoring.SubstitutionMatrix.SecondAlphabet.Encode(string(stringB[ep[0]]))
		return 0, e
	}
	return int(score[0]), nil
}

var _ = alphabet.DNA
var _ = fmt.Sprintf
