"""Independent Python big-int model of BLS12-377 used to pin the C oracle.

Nothing here shares code with oracle/ or the CUDA path: plain `int`, `pow(x, -1, p)`, affine
chord-and-tangent formulas and an O(n^2) DFT.  Decimal constants are the ones printed in the
reference's comments (arkworks/curves/bls12_377/src/fields/fr.rs:43,58,79, fq.rs:24,41,64,
curves/g1.rs:26,43-51, curves/g2.rs:26-32,68-86).
"""
import numpy as np

R_MOD = 8444461749428370424248824938781546531375899335154063827935233455917409239041
Q_MOD = 258664426012969094010652733694893533536393512754914660539884262666720468348340822774968888139573360124440321458177
FR_R_DEC = 6014086494747379908336260804527802945383293308637734276299549080986809532403
FQ_R_DEC = 85013442423176922659824578519796707547925331718418265885885478904210582549405549618995257669764901891699128663912
FR_GEN_MONT_DEC = 5642976643016801619665363617888466827793962762719196659561577942948671127251
FQ_GEN_MONT_DEC = 92261639910053574722182574790803529333160366917737991650341130812388023949653897454961487930322210790384999596794
FR_INV = 725501752471715839
FQ_INV = 9586122913090633727
FR_TWO_ADICITY = 47
FQ_TWO_ADICITY = 46
FR_T = (R_MOD - 1) >> FR_TWO_ADICITY
FQ_T = (Q_MOD - 1) >> FQ_TWO_ADICITY

G1_X = 81937999373150964239938255573465948239988671502647976594219695644855304257327692006745978603320413799295628339695
G1_Y = 241266749859715473739788878240585681733927191168601896383759122102112907357779751001206799952863815012735208165030
G1_COFACTOR = 30631250834960419227450344600217059328
G2_B = (0, 155198655607781456406391640216936120121836107652948796323930557600032281009004493664981332883744016074664192874906)
G2_X = (233578398248691099356572568220835526895379068987715365179118596935057653620464273615301663571204657964920925606294,
        140913150380207355837477652521042157274541796891053068589147167627541651775299824604154852141315666357241556069118)
G2_Y = (63160294768292073209381361943935198908131692476676907196754037919244929611450776219210369229519898517858833747423,
        149157405641012693445398062341192467754805999074082136895788947234480009303640899064710353187729182149407503257491)

FR_RR = 1 << 256
FQ_RR = 1 << 384


# ----------------------------------------------------------------------------- limb conversion
def to_limbs(v, n):
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def from_limbs(l):
    return sum(int(x) << (64 * i) for i, x in enumerate(l))


def fr_to_mont_arr(vals):
    """list of canonical ints -> (n,4) uint64 Montgomery array"""
    return np.array([to_limbs(v % R_MOD * FR_RR % R_MOD, 4) for v in vals], dtype=np.uint64).reshape(-1, 4)


def fr_from_mont_arr(arr):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    rinv = pow(FR_RR, -1, R_MOD)
    return [from_limbs(row) * rinv % R_MOD for row in arr]


def fq_to_mont_arr(vals):
    return np.array([to_limbs(v % Q_MOD * FQ_RR % Q_MOD, 6) for v in vals], dtype=np.uint64).reshape(-1, 6)


def fq_from_mont_arr(arr):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 6)
    rinv = pow(FQ_RR, -1, Q_MOD)
    return [from_limbs(row) * rinv % Q_MOD for row in arr]


# ----------------------------------------------------------------------------- Fq2 = Fq[u]/(u^2+5)
def fq2_mul(a, b):
    return ((a[0] * b[0] - 5 * a[1] * b[1]) % Q_MOD, (a[0] * b[1] + a[1] * b[0]) % Q_MOD)


def fq2_add(a, b):
    return ((a[0] + b[0]) % Q_MOD, (a[1] + b[1]) % Q_MOD)


def fq2_sub(a, b):
    return ((a[0] - b[0]) % Q_MOD, (a[1] - b[1]) % Q_MOD)


def fq2_inv(a):
    norm = (a[0] * a[0] + 5 * a[1] * a[1]) % Q_MOD
    ni = pow(norm, -1, Q_MOD)
    return (a[0] * ni % Q_MOD, (-a[1]) * ni % Q_MOD)


# ----------------------------------------------------------------------------- affine curve arithmetic
class Fp:
    """field adaptor so one set of curve formulas serves G1 (ints) and G2 (pairs)"""

    def __init__(self, ext):
        self.ext = ext
        self.zero = (0, 0) if ext else 0

    def add(self, a, b):
        return fq2_add(a, b) if self.ext else (a + b) % Q_MOD

    def sub(self, a, b):
        return fq2_sub(a, b) if self.ext else (a - b) % Q_MOD

    def mul(self, a, b):
        return fq2_mul(a, b) if self.ext else a * b % Q_MOD

    def inv(self, a):
        return fq2_inv(a) if self.ext else pow(a, -1, Q_MOD)

    def small(self, k, a):
        return ((k * a[0]) % Q_MOD, (k * a[1]) % Q_MOD) if self.ext else k * a % Q_MOD


F1, F2 = Fp(False), Fp(True)


def ec_add(F, P, Q):
    """P, Q are None (infinity) or (x, y); y^2 = x^3 + b, a = 0"""
    if P is None:
        return Q
    if Q is None:
        return P
    (x1, y1), (x2, y2) = P, Q
    if x1 == x2:
        if F.add(y1, y2) == F.zero:
            return None
        lam = F.mul(F.small(3, F.mul(x1, x1)), F.inv(F.small(2, y1)))
    else:
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
    y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
    return (x3, y3)


def ec_mul(F, k, P):
    acc = None
    while k:
        if k & 1:
            acc = ec_add(F, acc, P)
        P = ec_add(F, P, P)
        k >>= 1
    return acc


def ec_msm(F, points, scalars):
    acc = None
    for P, s in zip(points, scalars):
        acc = ec_add(F, acc, ec_mul(F, s % R_MOD, P))
    return acc


def g1_to_arr(P):
    """affine point (or None) -> ((12,) uint64 Montgomery x|y, inf flag) using the reference's
    affine zero (0, 1, infinity=true) (short_weierstrass_jacobian.rs:167-169)"""
    if P is None:
        return np.concatenate([fq_to_mont_arr([0])[0], fq_to_mont_arr([1])[0]]), 1
    return np.concatenate([fq_to_mont_arr([P[0]])[0], fq_to_mont_arr([P[1]])[0]]), 0


def g1_from_arr(xy, inf=0):
    if inf:
        return None
    v = fq_from_mont_arr(np.asarray(xy, dtype=np.uint64).reshape(2, 6))
    return (v[0], v[1])


def g2_to_arr(P):
    if P is None:
        return np.concatenate([fq_to_mont_arr([0, 0]).ravel(), fq_to_mont_arr([1, 0]).ravel()]), 1
    (x0, x1), (y0, y1) = P
    return fq_to_mont_arr([x0, x1, y0, y1]).ravel(), 0


def g2_from_arr(xy, inf=0):
    if inf:
        return None
    v = fq_from_mont_arr(np.asarray(xy, dtype=np.uint64).reshape(4, 6))
    return ((v[0], v[1]), (v[2], v[3]))


# ----------------------------------------------------------------------------- domains / DFT
def fr_root_of_unity(log_n):
    """get_root_of_unity: 22^T squared (47 - log_n) times"""
    w = pow(22, FR_T, R_MOD)
    for _ in range(FR_TWO_ADICITY - log_n):
        w = w * w % R_MOD
    return w


def dft(vals, kind):
    """O(n^2) in-order transform with the reference's semantics."""
    n = len(vals)
    log_n = n.bit_length() - 1
    w = fr_root_of_unity(log_n)
    g = 22
    if kind == "fft":
        return [sum(vals[j] * pow(w, i * j, R_MOD) for j in range(n)) % R_MOD for i in range(n)]
    if kind == "coset_fft":
        sh = [vals[j] * pow(g, j, R_MOD) % R_MOD for j in range(n)]
        return dft(sh, "fft")
    wi = pow(w, -1, R_MOD)
    ninv = pow(n, -1, R_MOD)
    out = [sum(vals[j] * pow(wi, i * j, R_MOD) for j in range(n)) * ninv % R_MOD for i in range(n)]
    if kind == "ifft":
        return out
    gi = pow(g, -1, R_MOD)
    return [out[i] * pow(gi, i, R_MOD) % R_MOD for i in range(n)]


def mix64(z):
    m = 0xFFFFFFFFFFFFFFFF
    z &= m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    return z ^ (z >> 31)


def gen_scalar_k(seed, i):
    k = mix64(seed + (i + 1) * 0x9E3779B97F4A7C15)
    return k if k else 1
