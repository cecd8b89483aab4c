// Package polyb200 is the cgo binding of libpolyb200.so (include/poly_b200.h): the thin
// shim the mash / align / primers drop-in packages call.  NOT COMPILED IN THIS REPOSITORY'S
// CI: the build image has no Go toolchain (DESIGN.md); the same C ABI is exercised by the
// C++ and Python host mirrors.  Build with:
//
//	CGO_CFLAGS="-I${POLYB200}/include" CGO_LDFLAGS="-L${POLYB200}/poly_b200/lib -lpolyb200" go build ./...
package polyb200

/*
#include <stdlib.h>
#include "poly_b200.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"runtime"
	"unsafe"
)

// ErrPanic is returned where the pure-Go reference would panic (index out of range).
var ErrPanic = errors.New("poly: reference panics on this input (index out of range)")

// locked runs one C ABI call with the goroutine pinned to its OS thread, so that pg_last_error()
// (a thread-local string) is read on the thread that produced it.
func locked(call func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	return check(call())
}

func check(rc C.int) error {
	switch rc {
	case C.PG_OK:
		return nil
	case C.PG_ERR_PANIC:
		return ErrPanic
	default:
		return fmt.Errorf("libpolyb200 error %d: %s", int(rc), C.GoString(C.pg_last_error()))
	}
}

// Flatten copies []string into the bytes+offsets layout of the C ABI.  The copy is what
// makes the call legal under the cgo pointer rules (no Go pointer to Go pointers crosses;
// nothing is retained by C after return).
func Flatten(seqs []string) ([]byte, []uint64) {
	total := 0
	for _, s := range seqs {
		total += len(s)
	}
	bases := make([]byte, 0, total+1)
	offsets := make([]uint64, len(seqs)+1)
	for i, s := range seqs {
		bases = append(bases, s...)
		offsets[i+1] = uint64(len(bases))
	}
	if len(bases) == 0 {
		bases = bases[:1] // keep &bases[0] valid
	}
	return bases, offsets
}

// SketchBatch wraps pg_mash_sketch_batch: compact rows of rowStride words, count[i] informative words.
func SketchBatch(bases []byte, offsets []uint64, k, s int, rowStride int) (out []uint32, count []uint32, status []int32, err error) {
	n := len(offsets) - 1
	out = make([]uint32, n*rowStride+1)
	count = make([]uint32, n+1)
	status = make([]int32, n+1)
	runtime.LockOSThread()
	defer runtime.UnlockOSThread() // pg_last_error() is thread-local: read it on the calling thread
	rc := C.pg_mash_sketch_batch((*C.uint8_t)(unsafe.Pointer(&bases[0])), (*C.uint64_t)(unsafe.Pointer(&offsets[0])), C.uint64_t(n),
		C.int32_t(k), C.int32_t(s), 0, (*C.uint32_t)(unsafe.Pointer(&out[0])), C.uint64_t(rowStride),
		(*C.uint32_t)(unsafe.Pointer(&count[0])), (*C.int32_t)(unsafe.Pointer(&status[0])))
	return out, count, status, check(rc)
}

func devicePtr(devices []int32) (*C.int32_t, C.int32_t) {
	if len(devices) == 0 {
		return nil, 0 // all visible GPUs
	}
	return (*C.int32_t)(unsafe.Pointer(&devices[0])), C.int32_t(len(devices))
}

// SketchBatchMulti wraps pg_mash_sketch_batch_multi: the batch is cut into contiguous shards, one per
// GPU of this process (devices == nil: every visible GPU), each driven by its own host thread inside the
// library.  Same results as SketchBatch.
func SketchBatchMulti(bases []byte, offsets []uint64, k, s int, rowStride int, devices []int32) (out []uint32, count []uint32, status []int32, err error) {
	n := len(offsets) - 1
	out = make([]uint32, n*rowStride+1)
	count = make([]uint32, n+1)
	status = make([]int32, n+1)
	dp, dn := devicePtr(devices)
	err = locked(func() C.int {
		return C.pg_mash_sketch_batch_multi((*C.uint8_t)(unsafe.Pointer(&bases[0])), (*C.uint64_t)(unsafe.Pointer(&offsets[0])), C.uint64_t(n),
			C.int32_t(k), C.int32_t(s), 0, (*C.uint32_t)(unsafe.Pointer(&out[0])), C.uint64_t(rowStride),
			(*C.uint32_t)(unsafe.Pointer(&count[0])), (*C.int32_t)(unsafe.Pointer(&status[0])), dp, dn)
	})
	return out, count, status, err
}

// SketchInto wraps pg_mash_sketch_batch_multi with PG_SKETCH_TAIL_KEEP: sketches is the caller's n x s slab (the
// Sketches arrays of n Mash values back to back); row i receives its min(len_i-k, s) words and nothing else is
// written -- a fresh make([]uint32, n*s) therefore ends up as n full, zero-tailed Sketches arrays without a
// per-read allocation or copy on the Go side, and a reused slab keeps its tails as (*Mash).Sketch does.
func SketchInto(bases []byte, offsets []uint64, k, s int, sketches []uint32, devices []int32) (count []uint32, status []int32, err error) {
	n := len(offsets) - 1
	if len(sketches) < n*s {
		return nil, nil, fmt.Errorf("polyb200: sketches holds %d words, need %d", len(sketches), n*s)
	}
	count = make([]uint32, n+1)
	status = make([]int32, n+1)
	if n == 0 {
		return count[:n], status[:n], nil
	}
	if len(bases) == 0 {
		bases = make([]byte, 1)
	}
	if len(sketches) == 0 { // s == 0: the call still reports the reads on which the reference panics
		sketches = make([]uint32, 1)
	}
	dp, dn := devicePtr(devices)
	err = locked(func() C.int {
		return C.pg_mash_sketch_batch_multi((*C.uint8_t)(unsafe.Pointer(&bases[0])), (*C.uint64_t)(unsafe.Pointer(&offsets[0])), C.uint64_t(n),
			C.int32_t(k), C.int32_t(s), C.PG_SKETCH_TAIL_KEEP, (*C.uint32_t)(unsafe.Pointer(&sketches[0])), C.uint64_t(s),
			(*C.uint32_t)(unsafe.Pointer(&count[0])), (*C.int32_t)(unsafe.Pointer(&status[0])), dp, dn)
	})
	return count[:n], status[:n], err
}

// SketchDistanceMulti wraps pg_mash_sketch_distance_multi for fixed-length reads stored back to back:
// sketches (n x s full arrays), matching counts and distances (n x n), computed on all listed GPUs with
// the all-gather of the sketches fused into the sketch kernels.
func SketchDistanceMulti(bases []byte, n, readLen, k, s int, devices []int32) (sketches []uint32, same []uint32, dist []float64, err error) {
	sketches = make([]uint32, n*s+1)
	same = make([]uint32, n*n+1)
	dist = make([]float64, n*n+1)
	dp, dn := devicePtr(devices)
	err = locked(func() C.int {
		return C.pg_mash_sketch_distance_multi((*C.uint8_t)(unsafe.Pointer(&bases[0])), C.uint64_t(n), C.uint32_t(readLen), C.int32_t(k), C.int32_t(s),
			dp, dn, (*C.uint32_t)(unsafe.Pointer(&sketches[0])), (*C.uint32_t)(unsafe.Pointer(&same[0])), (*C.double)(unsafe.Pointer(&dist[0])))
	})
	return sketches[:n*s], same[:n*n], dist[:n*n], err
}

// SimilarityPairs wraps pg_mash_similarity_pairs.
func SimilarityPairs(sketches []uint32, skOffsets []uint64, pairA, pairB []uint32) (same []int64, sim, dist []float64, err error) {
	np := len(pairA)
	same, sim, dist = make([]int64, np+1), make([]float64, np+1), make([]float64, np+1)
	status := make([]int32, np+1)
	if len(sketches) == 0 {
		sketches = make([]uint32, 1)
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread() // pg_last_error() is thread-local: read it on the calling thread
	rc := C.pg_mash_similarity_pairs((*C.uint32_t)(unsafe.Pointer(&sketches[0])), (*C.uint64_t)(unsafe.Pointer(&skOffsets[0])),
		C.uint64_t(len(skOffsets)-1), (*C.uint32_t)(unsafe.Pointer(&pairA[0])), (*C.uint32_t)(unsafe.Pointer(&pairB[0])), C.uint64_t(np),
		(*C.int64_t)(unsafe.Pointer(&same[0])), (*C.double)(unsafe.Pointer(&sim[0])), (*C.double)(unsafe.Pointer(&dist[0])),
		(*C.int32_t)(unsafe.Pointer(&status[0])))
	return same[:np], sim[:np], dist[:np], check(rc)
}

// DistanceBlock wraps pg_mash_distance_block (rows [rowBegin,rowEnd) x all n columns).
func DistanceBlock(sketches []uint32, n, s int, rowBegin, rowEnd int) (same []uint32, dist []float64, err error) {
	rows := rowEnd - rowBegin
	same, dist = make([]uint32, rows*n+1), make([]float64, rows*n+1)
	runtime.LockOSThread()
	defer runtime.UnlockOSThread() // pg_last_error() is thread-local: read it on the calling thread
	rc := C.pg_mash_distance_block((*C.uint32_t)(unsafe.Pointer(&sketches[0])), C.uint64_t(n), C.int32_t(s), C.uint64_t(rowBegin),
		C.uint64_t(rowEnd), (*C.uint32_t)(unsafe.Pointer(&same[0])), (*C.double)(unsafe.Pointer(&dist[0])))
	return same[:rows*n], dist[:rows*n], check(rc)
}

// DistanceSparse wraps pg_mash_distance_sparse: the pairs (i, j) of rows [rowBegin,rowEnd) x all n columns that share
// at least one hash, with their matching counts; every pair not listed has Similarity 0 / Distance 1
// (mash.go:134,139).  upper: only j > i.  The order of the triples is unspecified.
func DistanceSparse(sketches []uint32, n, s int, rowBegin, rowEnd int, upper bool) (pi, pj, same []uint32, err error) {
	capPairs := 4*(rowEnd-rowBegin) + 1024
	flags := C.uint32_t(0)
	if upper {
		flags = C.PG_PAIRS_UPPER
	}
	for {
		pi, pj, same = make([]uint32, capPairs), make([]uint32, capPairs), make([]uint32, capPairs)
		var cnt C.uint64_t
		var rc C.int
		err = locked(func() C.int {
			rc = C.pg_mash_distance_sparse((*C.uint32_t)(unsafe.Pointer(&sketches[0])), C.uint64_t(n), C.int32_t(s), C.uint64_t(rowBegin),
				C.uint64_t(rowEnd), flags, (*C.uint32_t)(unsafe.Pointer(&pi[0])), (*C.uint32_t)(unsafe.Pointer(&pj[0])),
				(*C.uint32_t)(unsafe.Pointer(&same[0])), C.uint64_t(capPairs), &cnt)
			return rc
		})
		if rc == C.PG_ERR_ARG && int(cnt) > capPairs { // the buffers were too small: the call reports the size it needs
			capPairs = int(cnt)
			continue
		}
		if err != nil {
			return nil, nil, nil, err
		}
		k := int(cnt)
		return pi[:k], pj[:k], same[:k], nil
	}
}

// SWScoreBatch wraps pg_sw_score_batch.
func SWScoreBatch(queries []byte, qOffsets []uint64, template string, queryIsA bool, lutA, lutB *[256]int16, table []int64,
	nA, nB int, gap int64) (score []int64, errCode []int32, errPos []int64, err error) {
	n := len(qOffsets) - 1
	score, errCode, errPos = make([]int64, n+1), make([]int32, n+1), make([]int64, n+1)
	t := []byte(template)
	if len(t) == 0 {
		t = make([]byte, 1)
	}
	qa := C.int32_t(0)
	if queryIsA {
		qa = 1
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread() // pg_last_error() is thread-local: read it on the calling thread
	rc := C.pg_sw_score_batch((*C.uint8_t)(unsafe.Pointer(&queries[0])), (*C.uint64_t)(unsafe.Pointer(&qOffsets[0])), C.uint64_t(n),
		(*C.uint8_t)(unsafe.Pointer(&t[0])), C.uint64_t(len(template)), qa, (*C.int16_t)(unsafe.Pointer(&lutA[0])),
		(*C.int16_t)(unsafe.Pointer(&lutB[0])), (*C.int64_t)(unsafe.Pointer(&table[0])), C.int32_t(nA), C.int32_t(nB), C.int64_t(gap),
		(*C.int64_t)(unsafe.Pointer(&score[0])), (*C.int32_t)(unsafe.Pointer(&errCode[0])), (*C.int64_t)(unsafe.Pointer(&errPos[0])))
	return score[:n], errCode[:n], errPos[:n], check(rc)
}

// NWScoreBatch wraps pg_nw_score_batch (same arguments as SWScoreBatch).
func NWScoreBatch(queries []byte, qOffsets []uint64, template string, queryIsA bool, lutA, lutB *[256]int16, table []int64,
	nA, nB int, gap int64) (score []int64, errCode []int32, errPos []int64, err error) {
	n := len(qOffsets) - 1
	score, errCode, errPos = make([]int64, n+1), make([]int32, n+1), make([]int64, n+1)
	t := []byte(template)
	if len(t) == 0 {
		t = make([]byte, 1)
	}
	qa := C.int32_t(0)
	if queryIsA {
		qa = 1
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread() // pg_last_error() is thread-local: read it on the calling thread
	rc := C.pg_nw_score_batch((*C.uint8_t)(unsafe.Pointer(&queries[0])), (*C.uint64_t)(unsafe.Pointer(&qOffsets[0])), C.uint64_t(n),
		(*C.uint8_t)(unsafe.Pointer(&t[0])), C.uint64_t(len(template)), qa, (*C.int16_t)(unsafe.Pointer(&lutA[0])),
		(*C.int16_t)(unsafe.Pointer(&lutB[0])), (*C.int64_t)(unsafe.Pointer(&table[0])), C.int32_t(nA), C.int32_t(nB), C.int64_t(gap),
		(*C.int64_t)(unsafe.Pointer(&score[0])), (*C.int32_t)(unsafe.Pointer(&errCode[0])), (*C.int64_t)(unsafe.Pointer(&errPos[0])))
	return score[:n], errCode[:n], errPos[:n], check(rc)
}

// NWAlignBatch wraps pg_nw_align_batch (same contract as SWAlignBatch).
func NWAlignBatch(queries []byte, qOffsets []uint64, template string, queryIsA bool, lutA, lutB *[256]int16, table []int64,
	nA, nB int, gap int64, stride int) ([]int64, []int32, []int64, []string, []string, error) {
	return alignBatch(true, queries, qOffsets, template, queryIsA, lutA, lutB, table, nA, nB, gap, stride)
}

// SWAlignBatch wraps pg_sw_align_batch: scores plus the two aligned strings per query.  Rows that
// do not fit `stride` bytes are retried with the reported length.
func SWAlignBatch(queries []byte, qOffsets []uint64, template string, queryIsA bool, lutA, lutB *[256]int16, table []int64,
	nA, nB int, gap int64, stride int) ([]int64, []int32, []int64, []string, []string, error) {
	return alignBatch(false, queries, qOffsets, template, queryIsA, lutA, lutB, table, nA, nB, gap, stride)
}

func alignBatch(global bool, queries []byte, qOffsets []uint64, template string, queryIsA bool, lutA, lutB *[256]int16, table []int64,
	nA, nB int, gap int64, stride int) (score []int64, errCode []int32, errPos []int64, alignA, alignB []string, err error) {
	n := len(qOffsets) - 1
	t := []byte(template)
	if len(t) == 0 {
		t = make([]byte, 1)
	}
	qa := C.int32_t(0)
	if queryIsA {
		qa = 1
	}
	for {
		score, errCode, errPos = make([]int64, n+1), make([]int32, n+1), make([]int64, n+1)
		oa, ob := make([]byte, n*stride+1), make([]byte, n*stride+1)
		ln, st := make([]uint32, n+1), make([]int32, n+1)
		var rc C.int
		if global {
			rc = C.pg_nw_align_batch((*C.uint8_t)(unsafe.Pointer(&queries[0])), (*C.uint64_t)(unsafe.Pointer(&qOffsets[0])), C.uint64_t(n),
				(*C.uint8_t)(unsafe.Pointer(&t[0])), C.uint64_t(len(template)), qa, (*C.int16_t)(unsafe.Pointer(&lutA[0])),
				(*C.int16_t)(unsafe.Pointer(&lutB[0])), (*C.int64_t)(unsafe.Pointer(&table[0])), C.int32_t(nA), C.int32_t(nB), C.int64_t(gap),
				(*C.int64_t)(unsafe.Pointer(&score[0])), (*C.int32_t)(unsafe.Pointer(&errCode[0])), (*C.int64_t)(unsafe.Pointer(&errPos[0])),
				(*C.uint8_t)(unsafe.Pointer(&oa[0])), (*C.uint8_t)(unsafe.Pointer(&ob[0])), C.uint64_t(stride),
				(*C.uint32_t)(unsafe.Pointer(&ln[0])), (*C.int32_t)(unsafe.Pointer(&st[0])))
		} else {
			rc = C.pg_sw_align_batch((*C.uint8_t)(unsafe.Pointer(&queries[0])), (*C.uint64_t)(unsafe.Pointer(&qOffsets[0])), C.uint64_t(n),
				(*C.uint8_t)(unsafe.Pointer(&t[0])), C.uint64_t(len(template)), qa, (*C.int16_t)(unsafe.Pointer(&lutA[0])),
				(*C.int16_t)(unsafe.Pointer(&lutB[0])), (*C.int64_t)(unsafe.Pointer(&table[0])), C.int32_t(nA), C.int32_t(nB), C.int64_t(gap),
				(*C.int64_t)(unsafe.Pointer(&score[0])), (*C.int32_t)(unsafe.Pointer(&errCode[0])), (*C.int64_t)(unsafe.Pointer(&errPos[0])),
				(*C.uint8_t)(unsafe.Pointer(&oa[0])), (*C.uint8_t)(unsafe.Pointer(&ob[0])), C.uint64_t(stride),
				(*C.uint32_t)(unsafe.Pointer(&ln[0])), (*C.int32_t)(unsafe.Pointer(&st[0])))
		}
		if e := check(rc); e != nil {
			return nil, nil, nil, nil, nil, e
		}
		grow := 0
		for i := 0; i < n; i++ {
			if st[i] == 2 && int(ln[i]) > grow {
				grow = int(ln[i])
			}
		}
		if grow > 0 {
			stride = grow + 8
			continue
		}
		alignA, alignB = make([]string, n), make([]string, n)
		for i := 0; i < n; i++ {
			alignA[i] = string(oa[i*stride : i*stride+int(ln[i])])
			alignB[i] = string(ob[i*stride : i*stride+int(ln[i])])
		}
		return score[:n], errCode[:n], errPos[:n], alignA, alignB, nil
	}
}

// TmBatch wraps pg_tm_batch.
func TmBatch(bases []byte, offsets []uint64, cp, na, mg float64) (tm, dH, dS []float64, status []int32, err error) {
	n := len(offsets) - 1
	tm, dH, dS, status = make([]float64, n+1), make([]float64, n+1), make([]float64, n+1), make([]int32, n+1)
	runtime.LockOSThread()
	defer runtime.UnlockOSThread() // pg_last_error() is thread-local: read it on the calling thread
	rc := C.pg_tm_batch((*C.uint8_t)(unsafe.Pointer(&bases[0])), (*C.uint64_t)(unsafe.Pointer(&offsets[0])), C.uint64_t(n),
		C.double(cp), C.double(na), C.double(mg), (*C.double)(unsafe.Pointer(&tm[0])), (*C.double)(unsafe.Pointer(&dH[0])),
		(*C.double)(unsafe.Pointer(&dS[0])), (*C.int32_t)(unsafe.Pointer(&status[0])))
	return tm[:n], dH[:n], dS[:n], status[:n], check(rc)
}

// FastaIngest parses a whole FASTA text on the GPU (pg_fasta_ingest): dense sequence bytes +
// offsets and dense name bytes + offsets for the records fasta.Parse would return, and the
// error it would stop with (errCode 0 = nil; errLine = the line number its message prints).
func FastaIngest(text []byte, maxLineSize int, bufioAlias bool) (seq []byte, seqOff []uint64, names []byte, nameOff []uint64, errCode int, errLine uint64, err error) {
	if len(text) == 0 {
		return nil, []uint64{0}, nil, []uint64{0}, 0, 0, nil
	}
	lines := 1
	for _, c := range text {
		if c == '\n' {
			lines++
		}
	}
	seq, names = make([]byte, len(text)), make([]byte, len(text))
	seqOff, nameOff = make([]uint64, lines+1), make([]uint64, lines+1)
	var n, tot, ntot C.uint64_t
	var ec C.int32_t
	var el C.uint64_t
	flags := C.uint32_t(0)
	if bufioAlias {
		flags = C.PG_FASTA_BUFIO_ALIAS
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread() // pg_last_error() is thread-local: read it on the calling thread
	rc := C.pg_fasta_ingest((*C.uint8_t)(unsafe.Pointer(&text[0])), C.uint64_t(len(text)), C.uint32_t(maxLineSize), flags,
		(*C.uint8_t)(unsafe.Pointer(&seq[0])), C.uint64_t(len(seq)), (*C.uint64_t)(unsafe.Pointer(&seqOff[0])),
		(*C.uint8_t)(unsafe.Pointer(&names[0])), C.uint64_t(len(names)), (*C.uint64_t)(unsafe.Pointer(&nameOff[0])),
		C.uint64_t(lines), &n, &tot, &ntot, &ec, &el)
	if err = check(rc); err != nil {
		return nil, nil, nil, nil, 0, 0, err
	}
	return seq[:tot], seqOff[:n+1], names[:ntot], nameOff[:n+1], int(ec), uint64(el), nil
}

// DesignPrimersBatch wraps pg_design_primers_batch: per sequence the lengths of the forward and
// reverse primers of pcr.DesignPrimersWithOverhangs (status 1: the reference slices out of range).
func DesignPrimersBatch(bases []byte, offsets []uint64, targetTm float64) (fwdLen, revLen []uint32, status []int32, err error) {
	n := len(offsets) - 1
	fwdLen, revLen, status = make([]uint32, n+1), make([]uint32, n+1), make([]int32, n+1)
	base := (*C.uint8_t)(nil)
	if len(bases) > 0 {
		base = (*C.uint8_t)(unsafe.Pointer(&bases[0]))
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread() // pg_last_error() is thread-local: read it on the calling thread
	rc := C.pg_design_primers_batch(base, (*C.uint64_t)(unsafe.Pointer(&offsets[0])), C.uint64_t(n), C.double(targetTm),
		(*C.uint32_t)(unsafe.Pointer(&fwdLen[0])), (*C.uint32_t)(unsafe.Pointer(&revLen[0])), (*C.int32_t)(unsafe.Pointer(&status[0])))
	return fwdLen[:n], revLen[:n], status[:n], check(rc)
}

// MinimalPrimerBatch wraps pg_pcr_minimal_primer_batch (the loop of pcr.go:93-100 for every primer).
func MinimalPrimerBatch(bases []byte, offsets []uint64, targetTm float64) (minLen []uint32, status []int32, err error) {
	n := len(offsets) - 1
	minLen, status = make([]uint32, n+1), make([]int32, n+1)
	base := (*C.uint8_t)(nil)
	if len(bases) > 0 {
		base = (*C.uint8_t)(unsafe.Pointer(&bases[0]))
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread() // pg_last_error() is thread-local: read it on the calling thread
	rc := C.pg_pcr_minimal_primer_batch(base, (*C.uint64_t)(unsafe.Pointer(&offsets[0])), C.uint64_t(n), C.double(targetTm),
		(*C.uint32_t)(unsafe.Pointer(&minLen[0])), (*C.int32_t)(unsafe.Pointer(&status[0])))
	return minLen[:n], status[:n], check(rc)
}

// Site is one exact occurrence of a pattern in a sequence.
type Site struct {
	Seq     int
	Pos     int
	Pattern int
}

// FindSites wraps pg_find_sites_batch: every (overlapping) occurrence of every pattern in every
// sequence, what suffixarray.Lookup(pattern, -1) returns per sequence.  Order is unspecified.
func FindSites(seqs []byte, seqOff []uint64, patterns []byte, patOff []uint64) ([]Site, error) {
	nSeq, nPat := len(seqOff)-1, len(patOff)-1
	if nSeq <= 0 || nPat <= 0 || len(seqs) == 0 || len(patterns) == 0 {
		return nil, nil
	}
	capHits := 1024
	for {
		hs, hp, hq := make([]uint32, capHits), make([]uint64, capHits), make([]uint32, capHits)
		var n C.uint64_t
		runtime.LockOSThread()
		defer runtime.UnlockOSThread() // pg_last_error() is thread-local: read it on the calling thread
		rc := C.pg_find_sites_batch((*C.uint8_t)(unsafe.Pointer(&seqs[0])), (*C.uint64_t)(unsafe.Pointer(&seqOff[0])), C.uint64_t(nSeq),
			(*C.uint8_t)(unsafe.Pointer(&patterns[0])), (*C.uint64_t)(unsafe.Pointer(&patOff[0])), C.uint32_t(nPat), 0,
			(*C.uint32_t)(unsafe.Pointer(&hs[0])), (*C.uint64_t)(unsafe.Pointer(&hp[0])), (*C.uint32_t)(unsafe.Pointer(&hq[0])),
			C.uint64_t(capHits), &n)
		if rc == C.PG_ERR_ARG && int(n) > capHits {
			capHits = int(n)
			continue
		}
		if err := check(rc); err != nil {
			return nil, err
		}
		sites := make([]Site, int(n))
		for i := range sites {
			sites[i] = Site{Seq: int(hs[i]), Pos: int(hp[i]), Pattern: int(hq[i])}
		}
		return sites, nil
	}
}

// FastqIngestRecords wraps pg_fastq_ingest_records: dense sequences + offsets, and per record the
// spans {identifier line begin, length, quality line begin, length} into text.
func FastqIngestRecords(text []byte) (bases []byte, offsets []uint64, spans []uint64, errCode int32, errLine uint64, err error) {
	newlines := 0
	for _, c := range text {
		if c == '\n' {
			newlines++
		}
	}
	capRec := newlines/4 + 1
	bases = make([]byte, len(text)+1)
	offsets = make([]uint64, capRec+1)
	spans = make([]uint64, 4*(capRec+1))
	var n, total C.uint64_t
	var code C.int32_t
	var line C.uint64_t
	var tp *C.uint8_t
	if len(text) > 0 {
		tp = (*C.uint8_t)(unsafe.Pointer(&text[0]))
	}
	err = locked(func() C.int {
		return C.pg_fastq_ingest_records(tp, C.uint64_t(len(text)), (*C.uint8_t)(unsafe.Pointer(&bases[0])), C.uint64_t(len(bases)),
			(*C.uint64_t)(unsafe.Pointer(&offsets[0])), (*C.uint64_t)(unsafe.Pointer(&spans[0])), C.uint64_t(capRec), &n, &total, &code, &line)
	})
	return bases[:total], offsets[:n+1], spans[:4*n], int32(code), uint64(line), err
}
