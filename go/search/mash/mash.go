// Package mash is the drop-in for github.com/bebop/poly/search/mash backed by libpolyb200.so:
// same exported names and semantics (search/mash/mash.go:52-140 of the reference), plus
// batched entry points.  NOT COMPILED HERE (no Go toolchain in the build image).
package mash

import (
	"errors"

	"github.com/bebop/poly/internal/polyb200"
)

// Mash is a collection of hashes of kmers from a given sequence (mash.go:52-56).
type Mash struct {
	KmerSize   int
	SketchSize int
	Sketches   []uint32
}

// New initializes a new mash sketch (mash.go:59-65).
func New(kmerSize int, sketchSize int) *Mash {
	return &Mash{KmerSize: kmerSize, SketchSize: sketchSize, Sketches: make([]uint32, sketchSize)}
}

// Sketch generates a mash sketch of the sequence (mash.go:68-104), on the GPU.
func (mash *Mash) Sketch(sequence string) {
	n := len(sequence) - mash.KmerSize
	if n <= 0 {
		return // the reference's loop body never runs (mash.go:73)
	}
	cnt := n
	if cnt > mash.SketchSize {
		cnt = mash.SketchSize
	}
	bases, offsets := polyb200.Flatten([]string{sequence})
	stride := cnt
	if stride < 1 {
		stride = 1
	}
	out, _, _, err := polyb200.SketchBatch(bases, offsets, mash.KmerSize, mash.SketchSize, stride)
	if errors.Is(err, polyb200.ErrPanic) {
		panic("runtime error: index out of range [-1]") // mash.go:96-98 with sketchSize in {0,1}
	} else if err != nil {
		panic(err)
	}
	copy(mash.Sketches[:cnt], out[:cnt]) // n < s: the tail keeps its previous contents (mash.go:81-84)
}

func (mash *Mash) pair(other *Mash) (float64, float64) {
	sk := append(append([]uint32{}, mash.Sketches[:mash.SketchSize]...), other.Sketches[:other.SketchSize]...)
	off := []uint64{0, uint64(mash.SketchSize), uint64(mash.SketchSize + other.SketchSize)}
	_, sim, dist, err := polyb200.SimilarityPairs(sk, off, []uint32{0}, []uint32{1})
	if errors.Is(err, polyb200.ErrPanic) {
		panic("runtime error: index out of range [-1]")
	} else if err != nil {
		panic(err)
	}
	return sim[0], dist[0]
}

// Similarity returns the Jaccard similarity between two sketches (mash.go:107-135).
func (mash *Mash) Similarity(other *Mash) float64 { s, _ := mash.pair(other); return s }

// Distance returns the Jaccard distance between two sketches (mash.go:138-140).
func (mash *Mash) Distance(other *Mash) float64 { _, d := mash.pair(other); return d }

// SketchBatch is New(kmerSize, sketchSize) + Sketch(seq) for every sequence: one call, sharded over every
// GPU of the box (reads are independent, mash.go:68-104; no collective is involved).
func SketchBatch(sequences []string, kmerSize, sketchSize int) []*Mash {
	if sketchSize < 0 {
		panic("runtime error: makeslice: len out of range") // mash.New would
	}
	bases, offsets := polyb200.Flatten(sequences)
	// one zeroed slab for all Sketches arrays (the runtime hands out zeroed memory, as New's make does per read);
	// the library writes the informative words of every row in place and leaves the zero tails alone
	slab := make([]uint32, len(sequences)*sketchSize)
	_, status, err := polyb200.SketchInto(bases, offsets, kmerSize, sketchSize, slab, nil)
	if err != nil && !errors.Is(err, polyb200.ErrPanic) {
		panic(err)
	}
	res := make([]*Mash, len(sequences))
	for i := range sequences {
		if status[i] != 0 {
			panic("runtime error: index out of range [-1]")
		}
		res[i] = &Mash{KmerSize: kmerSize, SketchSize: sketchSize,
			Sketches: slab[i*sketchSize : (i+1)*sketchSize : (i+1)*sketchSize]}
	}
	return res
}

// DistanceMatrix returns D[i][j] = sketches[i].Distance(sketches[j]) for sketches of one common SketchSize.
func DistanceMatrix(sketches []*Mash) [][]float64 {
	n := len(sketches)
	if n == 0 {
		return nil
	}
	s := sketches[0].SketchSize
	flat := make([]uint32, 0, n*s)
	for _, m := range sketches {
		flat = append(flat, m.Sketches[:s]...)
	}
	_, dist, err := polyb200.DistanceBlock(flat, n, s, 0, n)
	if err != nil {
		panic(err)
	}
	res := make([][]float64, n)
	for i := range res {
		res[i] = dist[i*n : (i+1)*n]
	}
	return res
}

// RelatedPair is one pair of sketches that share at least one hash.
type RelatedPair struct {
	I, J     int     // indices into the sketch list, I < J
	Same     int     // matching hashes as the walk of Similarity counts them (mash.go:121-132)
	Distance float64 // sketches[I].Distance(sketches[J]) == 1 - Same/SketchSize (mash.go:134,139)
}

// RelatedPairs returns every pair i < j of ascending (select-regime) or zero-padded sketches of one common SketchSize
// whose Distance is below 1, i.e. the non-trivial entries of DistanceMatrix, without materialising the n x n matrix
// (for n = 100k sketches: 334 MB instead of 40 GB).  Pairs that are not returned have Distance 1.
func RelatedPairs(sketches []*Mash) []RelatedPair {
	n := len(sketches)
	if n == 0 {
		return nil
	}
	s := sketches[0].SketchSize
	flat := make([]uint32, 0, n*s)
	for _, m := range sketches {
		flat = append(flat, m.Sketches[:s]...)
	}
	pi, pj, same, err := polyb200.DistanceSparse(flat, n, s, 0, n, true)
	if err != nil {
		panic(err)
	}
	out := make([]RelatedPair, len(pi))
	for t := range pi {
		out[t] = RelatedPair{I: int(pi[t]), J: int(pj[t]), Same: int(same[t]), Distance: 1 - float64(same[t])/float64(s)}
	}
	return out
}
