"""Numerical check (CPU, binary64) of the chord bound behind the packet filter (RTC_FLAG_PACKET, rtc_trace.cu: test_packet_groups):
if some ray of an 8-ray (4-ray) packet reaches |h| >= sqrt(kappa - eps0), an END ray reaches sqrt(kappa - E), E = 1e-5 + 1.5 D^2.
Random packets of several frame sizes, spheres placed on and next to packet rays (sub-pixel to large).  Prints violations
(must be 0) -- and, for the second part, how small the D^2 coefficient can get before violations appear (the bound's slack).
python scripts/experiments/check_packet_bound.py"""
import numpy as np

rng = np.random.default_rng(1)


def packets(y, W, e1, e2, n_pk, kr, R):
    rows0 = rng.integers(0, max(1, y - kr), n_pk)
    cols = rng.integers(0, W, n_pk)
    vx = ((2 * cols - (W + 1)) / (W + 1)) * e1
    vy = ((y - 2 * (rows0[:, None] + np.arange(kr)[None, :])) / y) * e2
    w = R[:, 2][None, None, :] + vx[:, None, None] * R[:, 0][None, None, :] + vy[:, :, None] * R[:, 1][None, None, :]
    return w / np.linalg.norm(w, axis=2, keepdims=True)


def random_rotation():
    a, b, c, d = (lambda q: q / np.linalg.norm(q))(rng.normal(size=4))
    return np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                     [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                     [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])


def check(y, W, e1, e2, n_pk=20000, n_sph=400, kr=8, tiny=False, coeff=1.5, base=1e-5):
    D = (kr - 1) * 2 * e2 / y
    E, eps0 = base + coeff * D * D, 2.7e-6
    dirs = packets(y, W, e1, e2, n_pk, kr, random_rotation())
    viol = flagged = hits = 0
    for _ in range(n_sph):
        k = rng.integers(0, n_pk)
        dist = rng.uniform(5, 200)
        cen = (dirs[k, rng.integers(0, kr)] + rng.normal(size=3) * rng.choice([1e-4, 1e-3, 1e-2, 5e-2])) * dist
        rad = dist * rng.uniform(0, 2e-3) if tiny else rng.choice([0.0, dist * rng.uniform(0, 3e-3), rng.uniform(0, 9)])
        oc2 = cen @ cen
        cc = oc2 - rad * rad
        if cc <= 0:
            continue
        kappa = cc / oc2
        h = dirs @ (-cen / np.sqrt(oc2))
        hit = (np.abs(h) >= np.sqrt(max(kappa - eps0, 0))).any(axis=1)
        flag = np.ones(n_pk, bool) if kappa <= E else (np.abs(h[:, 0]) >= np.sqrt(kappa - E)) | (np.abs(h[:, -1]) >= np.sqrt(kappa - E))
        viol += int((hit & ~flag).sum()); flagged += int(flag.sum()); hits += int(hit.sum())
    print(f"y={y:5d} rays/packet={kr} E={E:.3e} coeff={coeff}: packets with a hit {hits}, flagged {flagged}, violations {viol}")
    return viol


def main():
    bad = 0
    for args in ((2160, 3840, 0.325, 0.577), (1080, 1920, 0.325, 0.577), (150, 399, 0.866, 0.577), (4320, 7680, 0.325, 0.577), (2160, 3840, 3.0, 5.0)):
        for tiny in (False, True):
            bad += check(*args, tiny=tiny)
            bad += check(*args, tiny=tiny, kr=4)
    print("violations with the shipped bound:", bad)
    print("slack: spheres centred between rows 3 and 4 of a packet, E = eps0 + coeff * D^2")
    for coeff in (0.0, 0.05, 0.1, 0.2, 0.5):
        check(2160, 3840, 0.325, 0.577, n_pk=20000, n_sph=2000, tiny=True, coeff=coeff, base=2.7e-6)


if __name__ == "__main__":
    main()
