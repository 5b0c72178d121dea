"""Dev helper (design (a)): numpy model of a two-pass, one-exchange inverse FFT for the correlate
kernels' sub-transforms, with the thread ownership, the exchange pattern and the twiddle recurrences the CUDA version
would use.  Checks itself against numpy.fft:  python tools/proto_two_pass.py

Length N = RA * RB (512 = 16 * 32, 640 = 20 * 32), NT = RA threads per transform, thread j owns the RB elements
e = j + RA * s (s < RB) at load time AND at store time, so global accesses of a warp stay contiguous runs of RA elements.

  The larger radix goes FIRST: the owned elements e = j + RA * s are exactly the inputs of butterfly j of a radix-RB
  pass with T = RA butterflies.
  pass 1 (radix RB, Ns = 1, T = RA):     thread j: v[s] = x[j + RA * s], DFT_RB, writes y[RB * j + q], q < RB.
  exchange through shared memory (the only one).
  pass 2 (radix RA, Ns = RB, T = RB):    butterfly t (t < RB) reads y[t + RB * r] * w_N^{t r} (r < RA), writes
      z[t + RB * q] (q < RA).  A thread takes RB / RA butterflies... for 16 x 32 that is two radix-16 butterflies per
      thread (t = j and t = j + 16): it then owns z[j + 16 * m] for all m < 32 -- the same residue class as at load.
"""
import numpy as np


def dft(v, sign):
    n = len(v)
    k = np.arange(n)
    return np.exp(sign * 2j * np.pi * np.outer(k, k) / n) @ v


def two_pass(x, RA, RB, sign=+1):
    N = RA * RB
    assert len(x) == N and RB % RA == 0
    per = RB // RA                                   # radix-RA butterflies per thread in pass 2
    smem = np.zeros(N, dtype=complex)
    out = np.zeros(N, dtype=complex)
    owned_load = {}
    for j in range(RA):                              # pass 1: one radix-RB butterfly per thread
        idx = j + RA * np.arange(RB)
        owned_load[j] = set(idx.tolist())
        smem[RB * j + np.arange(RB)] = dft(x[idx], sign)          # Ns = 1: no twiddles
    for j in range(RA):                              # pass 2: `per` radix-RA butterflies per thread
        mine = set()
        for u in range(per):
            t = j + RA * u
            # twiddles w_N^{t r}, r < RA, by recurrence from the base w_N^t (what the kernel would do per unit)
            base = np.exp(sign * 2j * np.pi * t / N)
            tw = np.ones(RA, dtype=complex)
            for r in range(1, RA):
                tw[r] = tw[r - 1] * base
            v = smem[t + RB * np.arange(RA)] * tw
            dst = t + RB * np.arange(RA)
            out[dst] = dft(v, sign)
            mine.update(dst.tolist())
        assert mine == owned_load[j], "thread must own the same residue class at load and store"
    return out


def main():
    rng = np.random.default_rng(0)
    for RA, RB in ((16, 32), (20, 40)):
        N = RA * RB
        x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        got = two_pass(x, RA, RB, +1)
        want = np.fft.ifft(x) * N
        err = np.max(np.abs(got - want)) / np.max(np.abs(want))
        print(f"N = {N} = {RA} x {RB}: max rel err {err:.2e}")
        assert err < 1e-12
    print("the column pass (N1 = 640 = 20 x 32) feeds a max reduction, not memory, so its store-side ownership is free: "
          "32 threads x radix 20 on e = j + 32 s, one exchange, then 20 radix-32 butterflies")

if __name__ == "__main__":
    main()
