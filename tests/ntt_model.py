"""Pure-Python model of the index algebra of csrc/ntt.cu (pass planning, tile layout, twiddle
indices, mirrored inverse twiddles, bit-reversed store, fused coset scaling, buffer ping-pong).  It
mirrors the kernel thread-for-thread with Python ints so the addressing can be validated on a
CPU-only box against the O(n^2) DFT; tests/test_ntt_model.py runs it, ntt.cu cites it.
Test scaffolding, not a product path."""
import pyref as P

R = P.R_MOD


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def plan_passes(log_n, max_deg):
    """list of (s0, deg) covering stages 0..log_n-1 in passes of at most max_deg stages, split evenly"""
    if log_n == 0:
        return []
    npass = (log_n + max_deg - 1) // max_deg
    base, extra = divmod(log_n, npass)
    out, s0 = [], 0
    for i in range(npass):
        d = base + (1 if i < extra else 0)
        out.append((s0, d))
        s0 += d
    return out


def twiddle(log_n, s, k, inverse):
    """stage-s twiddle with exponent index k of a size-2^log_n DIF, read from the FORWARD table of the
    domain l = log_n - s (tw[l][i] = w_l^i, i < 2^(l-1)).  Returns (swap_sub, value): for the inverse
    transform w^-k = -w^(half-k), the sign being folded into the order of the subtraction."""
    l = log_n - s
    half = 1 << (l - 1)
    w = P.fr_root_of_unity(l)
    if not inverse or k == 0:
        return False, pow(w, k, R)
    return True, pow(w, half - k, R)


def run(vals, kind, max_deg=3, log_c=1):
    n = len(vals)
    log_n = n.bit_length() - 1
    inverse = kind in ("ifft", "coset_ifft")
    g, gi, ninv = 22, pow(22, -1, R), pow(n, -1, R)
    passes = plan_passes(log_n, max_deg)
    data = list(vals)
    if not passes:      # n == 1
        return [data[0] * ninv % R] if inverse else data
    tmp = [None] * n
    for pi, (s0, deg) in enumerate(passes):
        first, last = pi == 0, pi == len(passes) - 1
        # buffer ping-pong: 1 pass: data->data (one tile per vector); else data->tmp, tmp->tmp.., tmp->data
        src = data if first else tmp
        dst = data if last else tmp
        T = n >> (s0 + deg)
        rows = 1 << deg
        if not last:
            C = min(1 << log_c, T)
            tiles = (1 << s0) * (T // C)
        else:
            assert T == 1
            C = min(1 << log_c, 1 << s0)
            tiles = (1 << s0) // C
        lc = C.bit_length() - 1
        E = rows * C
        staged = []
        for tile in range(tiles):
            def gidx(j, c):
                if not last:
                    hi, lo0 = divmod(tile, T // C)
                    return hi * (n >> s0) + j * T + lo0 * C + c
                hi_c = bitrev(tile * C, s0) + (bitrev(c, lc) << (s0 - lc))
                return hi_c * rows + j
            sm = [0] * E
            for j in range(rows):
                for c in range(C):
                    i = gidx(j, c)
                    x = src[i]
                    if first and kind == "coset_fft":
                        x = x * pow(g, i, R) % R
                    sm[j * C + c] = x
            for r in range(deg):
                half = 1 << (deg - 1 - r)
                for t in range(E // 2):
                    c, q = t % C, t // C
                    j0 = (q // half) * 2 * half + (q % half)
                    j1 = j0 + half
                    lo = 0 if last else (tile % (T // C)) * C + c
                    k = (q % half) * T + lo
                    swap, w = twiddle(log_n, s0 + r, k, inverse)
                    u, v = sm[j0 * C + c], sm[j1 * C + c]
                    sm[j0 * C + c] = (u + v) % R
                    sm[j1 * C + c] = ((v - u) if swap else (u - v)) % R * w % R
            for j in range(rows):
                for c in range(C):
                    x = sm[j * C + c]
                    if not last:
                        staged.append((gidx(j, c), x))
                    else:
                        dest = bitrev(j, deg) * (1 << s0) + tile * C + c
                        assert dest == bitrev(gidx(j, c), log_n)
                        if kind == "ifft":
                            x = x * ninv % R
                        elif kind == "coset_ifft":
                            x = x * ninv % R * pow(gi, dest, R) % R
                        staged.append((dest, x))
            if len(passes) > 1 or True:
                # a tile's stores may only touch dst; when src is dst (middle passes, or the single
                # pass) the tile must read exactly the positions it writes
                if src is dst:
                    reads = sorted(gidx(j, c) for j in range(rows) for c in range(C))
                    writes = sorted(p for p, _ in staged[-E:])
                    assert reads == writes, "in-place pass must be tile-closed"
        for p, x in staged:
            dst[p] = x
    return data
