// Package primers: drop-in for primers.SantaLucia / MeltingTemp (primers/primers.go:70-105,
// 121-128) backed by libpolyb200.so.  NOT COMPILED HERE.
package primers

import "github.com/bebop/poly/internal/polyb200"

// SantaLucia mirrors primers.SantaLucia (primers.go:70-105).
func SantaLucia(sequence string, primerConcentration, saltConcentration, magnesiumConcentration float64) (meltingTemp, dH, dS float64) {
	bases, offsets := polyb200.Flatten([]string{sequence})
	tm, h, s, status, err := polyb200.TmBatch(bases, offsets, primerConcentration, saltConcentration, magnesiumConcentration)
	if status[0] == 1 {
		panic("runtime error: index out of range [-1]") // primers.go:89
	}
	if err != nil {
		panic(err)
	}
	return tm[0], h[0], s[0]
}

// MeltingTemp mirrors primers.MeltingTemp (primers.go:121-128).
func MeltingTemp(sequence string) float64 {
	tm, _, _ := SantaLucia(sequence, 500e-9, 50e-3, 0.0)
	return tm
}

// MeltingTemps evaluates MeltingTemp for every sequence in one GPU pass.
func MeltingTemps(sequences []string) []float64 {
	bases, offsets := polyb200.Flatten(sequences)
	tm, _, _, status, err := polyb200.TmBatch(bases, offsets, 500e-9, 50e-3, 0.0)
	for _, st := range status {
		if st == 1 {
			panic("runtime error: index out of range [-1]")
		}
	}
	if err != nil {
		panic(err)
	}
	return tm
}
