// basecomp.cuh -- the complement of one sequence byte as ma_ug_seq applies it on the reverse strand (asm.c:224-233 comp_tab,
// asm.c:281 for bytes >= 128), as arithmetic instead of a table; host + device so that the CPU tier can check all 256 values
// against the reference (tests/hostsim/hit_host.cpp, tests/test_hitrules_cpu.py).
#pragma once

__host__ __device__ __forceinline__ unsigned char mab_comp_of(unsigned char c) // asm.c:224-233 comp_tab, bytes >= 128 -> 'N' (asm.c:281)
{
	if (c >= 128) return 'N';
	if (c == 96) return 64;
	const unsigned char u = c & 0xdf, lower = c & 0x20;          // letters only below
	if (u < 'A' || u > 'Z') return c;
	unsigned char r;
	switch (u) {
		case 'A': r = 'T'; break; case 'T': r = 'A'; break; case 'U': r = 'A'; break; case 'C': r = 'G'; break; case 'G': r = 'C'; break;
		case 'B': r = 'V'; break; case 'V': r = 'B'; break; case 'D': r = 'H'; break; case 'H': r = 'D'; break;
		case 'K': r = 'M'; break; case 'M': r = 'K'; break; case 'R': r = 'Y'; break; case 'Y': r = 'R'; break;
		default: r = u;
	}
	return r | lower;
}
