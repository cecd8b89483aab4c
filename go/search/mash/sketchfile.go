// Sketch persistence for the drop-in mash package (SURVEY.md 8f.4).  The reference has no storage
// format: a Mash is three exported fields (search/mash/mash.go:52-56) and already round-trips through
// encoding/json as {"KmerSize":k,"SketchSize":s,"Sketches":[...]} -- MarshalJSON is NOT overridden here,
// so single sketches stay interchangeable with the reference.  For sketch SETS (what SketchBatch returns
// and DistanceMatrix consumes) this file reads and writes the PGSKETCH v1 container that the Python and
// C++ mirrors use (poly_b200/sketchfile.py, hostcpp/poly_b200.hpp): little endian,
//
//	"PGSKETCH" | version 1 | KmerSize | SketchSize | flags (bit 0: dense) | n u64 | words u64 |
//	count[n] u32 (absent when dense) | informative words of sketch 0, 1, ... | CRC-32 (IEEE) of all before
//
// Only the informative words are stored: the zero tail of a fresh Mash is re-materialised on load.
// NOT COMPILED IN THIS REPOSITORY (no Go toolchain in the build image); see INTEGRATION.md.
package mash

import (
	"bytes"
	"encoding/binary"
	"errors"
	"hash/crc32"
	"io"
	"os"
)

const sketchMagic = "PGSKETCH"

// informative words of a sketch: everything up to the last non-zero word (a fill-regime sketch is its
// hashes in positional order followed by the zeros New put there)
func informative(m *Mash) []uint32 {
	end := len(m.Sketches)
	for end > 0 && m.Sketches[end-1] == 0 {
		end--
	}
	return m.Sketches[:end]
}

// WriteSketchSet writes sketches of one common (KmerSize, SketchSize).
func WriteSketchSet(w io.Writer, sketches []*Mash) error {
	if len(sketches) == 0 {
		return errors.New("mash: empty sketch set")
	}
	k, s := sketches[0].KmerSize, sketches[0].SketchSize
	rows := make([][]uint32, len(sketches))
	dense, words := true, uint64(0)
	for i, m := range sketches {
		if m.KmerSize != k || m.SketchSize != s {
			return errors.New("mash: a sketch set needs one common KmerSize and SketchSize")
		}
		rows[i] = informative(m)
		dense = dense && len(rows[i]) == s
		words += uint64(len(rows[i]))
	}
	var buf bytes.Buffer
	buf.WriteString(sketchMagic)
	flags := uint32(0)
	if dense {
		flags = 1
	}
	for _, v := range []uint32{1, uint32(k), uint32(s), flags} {
		_ = binary.Write(&buf, binary.LittleEndian, v)
	}
	_ = binary.Write(&buf, binary.LittleEndian, uint64(len(rows)))
	_ = binary.Write(&buf, binary.LittleEndian, words)
	if !dense {
		for _, r := range rows {
			_ = binary.Write(&buf, binary.LittleEndian, uint32(len(r)))
		}
	}
	for _, r := range rows {
		_ = binary.Write(&buf, binary.LittleEndian, r)
	}
	_ = binary.Write(&buf, binary.LittleEndian, crc32.ChecksumIEEE(buf.Bytes()))
	_, err := w.Write(buf.Bytes())
	return err
}

// ReadSketchSet is the inverse of WriteSketchSet.
func ReadSketchSet(r io.Reader) ([]*Mash, error) {
	blob, err := io.ReadAll(r)
	if err != nil {
		return nil, err
	}
	if len(blob) < 44 || string(blob[:8]) != sketchMagic {
		return nil, errors.New("mash: not a PGSKETCH file")
	}
	body, tail := blob[:len(blob)-4], blob[len(blob)-4:]
	if crc32.ChecksumIEEE(body) != binary.LittleEndian.Uint32(tail) {
		return nil, errors.New("mash: sketch file checksum mismatch")
	}
	u32 := func(at int) uint32 { return binary.LittleEndian.Uint32(blob[at:]) }
	if u32(8) != 1 {
		return nil, errors.New("mash: not a PGSKETCH v1 file")
	}
	k, s, flags := int(u32(12)), int(u32(16)), u32(20)
	n, words := binary.LittleEndian.Uint64(blob[24:]), binary.LittleEndian.Uint64(blob[32:])
	pos := uint64(40)
	counts := make([]uint32, n)
	total := uint64(0)
	for i := range counts {
		if flags&1 != 0 {
			counts[i] = uint32(s)
		} else {
			if pos+4 > uint64(len(body)) {
				return nil, errors.New("mash: truncated sketch file")
			}
			counts[i] = u32(int(pos))
			pos += 4
		}
		total += uint64(counts[i])
	}
	if total != words || pos+4*words != uint64(len(body)) {
		return nil, errors.New("mash: sketch file is inconsistent")
	}
	out := make([]*Mash, n)
	for i := range out {
		m := New(k, s)
		for j := uint32(0); j < counts[i]; j++ {
			m.Sketches[j] = u32(int(pos))
			pos += 4
		}
		out[i] = m
	}
	return out, nil
}

// SaveSketchSet / LoadSketchSet: the same through a file.
func SaveSketchSet(path string, sketches []*Mash) error {
	f, err := os.Create(path)
	if err != nil {
		return err
	}
	defer f.Close()
	return WriteSketchSet(f, sketches)
}

func LoadSketchSet(path string) ([]*Mash, error) {
	f, err := os.Open(path)
	if err != nil {
		return nil, err
	}
	defer f.Close()
	return ReadSketchSet(f)
}
