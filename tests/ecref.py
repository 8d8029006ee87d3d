"""Pure-Python big-int helpers for the tests (independent of oracle/ and of the product)."""
P = 2**251 + 17 * 2**192 + 1
BETA = 3141592653589793238462643383279502884197169399375105820974944592307816406665
# builtins/src/pedersen/constants.rs:6-29
PEDERSEN_P = [
    (2089986280348253421170679821480865132823066470938446095505822317253594081284, 1713931329540660377023406109199410414810705867260802078187082345529207694986),
    (996781205833008774514500082376783249102396023663454813447423147977397232763, 1668503676786377725805489344771023921079126552019160156920634619255970485781),
    (2251563274489750535117886426533222435294046428347329203627021249169616184184, 1798716007562728905295480679789526322175868328062420237419143593021674992973),
    (2138414695194151160943305727036575959195309218611738193261179310511854807447, 113410276730064486255102093846540133784865286929052426931474106396135072156),
    (2379962749567351885752724891227938183011949129833673362440656643086021394946, 776496453633298175483985398648758586525933812536653089401905292063708816422),
]
# builtins/src/utils.rs:147-150
EC_GENERATOR = (874739451078007766457464989774322083649278607533249481151382481072868806602, 152666792071518830868575557812948353041420400780739481342941381225525861407)


def ec_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    (x1, y1), (x2, y2) = a, b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = (3 * x1 * x1 + 1) * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return x3, (lam * (x1 - x3) - y1) % P


def ec_mul(k, pt):
    acc = None
    while k:
        if k & 1:
            acc = ec_add(acc, pt)
        pt = ec_add(pt, pt)
        k >>= 1
    return acc


def pedersen_hash(a, b):
    acc = PEDERSEN_P[0]
    for i, v in enumerate((a, b)):
        lo, hi = v & (2**248 - 1), v >> 248
        acc = ec_add(acc, ec_mul(lo, PEDERSEN_P[1 + 2 * i]))
        acc = ec_add(acc, ec_mul(hi, PEDERSEN_P[2 + 2 * i]))
    return acc[0]


def root_of_unity(n):
    return pow(3, (P - 1) // n, P)


def naive_dft(coeffs, offset=1):
    n = len(coeffs)
    w = root_of_unity(n)
    return [sum(c * pow(offset * pow(w, i, P), k, P) for k, c in enumerate(coeffs)) % P for i in range(n)]
