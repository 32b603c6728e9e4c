"""GPU version of the reference's error-profile study (varsens/profile/parameters.py:51-83, :130-147).

For every n in the reference's grid: b random g-function models a ~ U(0, 99)^k on ONE sample of (k, n); error = sum over
factors of (V_i - sens_i * var_y)^2 against the closed form V_i = 1/(3 (1+a_i)^2) (parameters.py:14-15); the row written
per n is (n, mean, sd, lower CI, upper CI, max) exactly as error-profile-dim*.csv.  Each model is one run of the fused
kernel (k <= 20) or of the two-phase GPU path (any k); the sample is never materialised.

    python tools/accuracy_profile.py --k 6 --b 30 [--out profiles/error-profile-dim6.csv]
"""
import argparse, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
from scipy.stats import t as student_t
import varsens_b200 as vb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=6)
    ap.add_argument("--b", type=int, default=30)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=None)
    ap.add_argument("--nmax", type=int, default=20480)
    args = ap.parse_args()
    random.seed(args.seed)
    rows = []
    for n in [5, 10, 20, 40, 80, 160, 320, 640, 1280, 2560, 5120, 10240, 20480]:
        if n > args.nmax:
            break
        s = vb.Sample(args.k, n, lambda x: x, verbose=False)                      # parameters.py:75
        errs = []
        for _ in range(args.b):
            model = [random.uniform(0, 99) for _ in range(args.k)]                # parameters.py:56
            v = vb.Varsens(vb.GFunction(model), sample=s, verbose=False)
            truth = 1.0 / (3.0 * (numpy.array(model) + 1.0) ** 2.0)
            errs.append(float(numpy.sum((truth - v.sens[:, 0] * v.var_y[0]) ** 2)))
        mu, sd = numpy.mean(errs), numpy.std(errs)
        se = sd / numpy.sqrt(args.b)
        q = student_t.isf(0.025, args.b - 1)
        rows.append((n, mu, sd, mu - se * q, mu + se * q, numpy.amax(errs)))
        print("n=%6d  mean squared error %.3e  sd %.3e  max %.3e" % (n, mu, sd, rows[-1][5]))
    if args.out:
        numpy.savetxt(args.out, numpy.array(rows), delimiter=",")


if __name__ == "__main__":
    main()
