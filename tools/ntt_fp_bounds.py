# bound propagation (units of q) for the FP64 radix-16 blocks; rho = q / 2^50, need every intermediate <= 8/rho
# modmul_fp (common.cuh) estimates the quotient from the rounded product: |r| <= (0.5 + 1.5 |a| 2^-52) q
def mm(a, rho):  # |modmul output| bound for |input| <= a (units of q)
    return 0.5 + a * rho * 0.375 * 1.001 + 1e-9
def fwd(rho, fold_at=()):
    x = [0.51] * 16; worst = 0
    for s in range(4):
        span = 8 >> s
        if s in fold_at: x = [0.51] * 16
        for i in range(8):
            g = i // span; k0 = g * 2 * span + i % span; k1 = k0 + span
            u, v = x[k0], mm(x[k1], rho)
            worst = max(worst, x[k1], u + v)
            x[k0] = x[k1] = u + v
    return worst, max(x)
def inv(rho, folds=None, final=False):
    x = [0.51] * 16; worst = 0
    for s in range(4):
        span = 1 << s
        if folds and s in folds:
            for k in folds[s]: x[k] = 0.51
        for i in range(8):
            g = i // span; k0 = g * 2 * span + i % span; k1 = k0 + span
            u, v = x[k0], x[k1]
            worst = max(worst, u + v)
            x[k0] = mm(u + v, rho) if (final and s == 3) else u + v
            x[k1] = mm(u + v, rho)
    return worst, max(x), [round(t, 2) for t in x]
for rho in (1.0, 1.05, 1.1):
    print("rho", rho, "limit", 8 / rho)
    print("  fwd nofold", fwd(rho), " fold@2", fwd(rho, (2,)))
    print("  inv nofold", inv(rho)[:2])
    print("  inv fold {3:[0,8,1,9]}", inv(rho, {3: [0, 8, 1, 9]}))
    print("  inv fold {3:[0,8]}", inv(rho, {3: [0, 8]})[:2])
    print("  inv fold {2:[0,4]}", inv(rho, {2: [0, 4]})[:2])
    print("  inv fold {2:all}", inv(rho, {2: list(range(16))})[:2])
