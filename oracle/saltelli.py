"""Python-3/numpy restatement of varsens/saltelli.py.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows, in semantics, varsens/saltelli.py:62-125 (Sample), :127-160 (flat layout), :173-250
(text export/load), :287-355 (Objective evaluation order), :357-495 (Objective flat / export /
load / NaN trimming) and :545-622 (Varsens + estimators).  The *intended* API is the default
(SURVEY.md App. B): ``Varsens(callable, scaling, k, n, verbose=...)`` builds
``Sample(k, n, scaling, discard=0, verbose=verbose)`` and ``Objective(k, n, sample, callable,
verbose=verbose)``; reference HEAD passes ``verbose`` positionally into ``discard`` (:555) and
``objective_vals`` (:567, a TypeError).  ``literal_head=True`` reproduces the :555 slip.

The Halton source is oracle.halton (ghalton restatement, parity unpinned); the row shuffle is
numpy's own legacy global RNG exactly as the reference calls it (:100-101).
"""
import os

import numpy

from . import halton as _halton


def _say(verbose, *a):
    if verbose:
        print(*a)


class Sample(object):
    def __init__(self, k, n, scaling=None, discard=0, verbose=True, raw=None, **loadArgs):
        self.k, self.n = int(k), int(n)          # saltelli.py:64-65
        self.scaling, self.verbose = scaling, verbose
        if raw is not None:                      # :69-73
            x = raw
            if x.shape != (2 * self.n, self.k):
                raise Exception("Raw sample dimensions do not match specified dimensions")
        elif loadArgs:                           # :74-77
            x = self.load(**loadArgs)
            if x.shape == (2 * self.n * (1 + self.k), self.k):
                return
        else:                                    # :78-84
            if not self.scaling:
                raise Exception("Generating a fresh sample space requires that a 'scaling' function be defined.")
            seq = _halton.Halton(self.k)
            seq.get(20 * self.k + int(discard))
            x = numpy.array(seq.get(2 * self.n))
        if scaling is None:                      # :86-88
            self.scaling = lambda p: p
        self.M_1 = self.scaling(x[0:self.n, ...])            # :92
        self.M_2 = self.scaling(x[self.n:2 * self.n, ...])   # :95
        numpy.random.seed(1)                                 # :100
        numpy.random.shuffle(self.M_2)                       # :101
        self.N_j = self.generate_N_j(self.M_1, self.M_2)     # :105
        self.N_nj = self.generate_N_j(self.M_2, self.M_1)    # :108

    def generate_N_j(self, M_1, M_2):
        # :119-123 -- k copies of M_2, column i of copy i taken from M_1
        N = numpy.array([M_2] * self.k)
        for i in range(self.k):
            N[i, :, i] = M_1[:, i]
        return N

    def flat(self):
        # :127-160 -- rows: M_1, M_2, N_j[0..k), N_nj[0..k)
        n, k = self.n, self.k
        x = numpy.zeros((2 * n * (1 + k), k))
        x[0:n] = self.M_1
        x[n:2 * n] = self.M_2
        at = 2 * n
        for blockset in (self.N_j, self.N_nj):
            for i in range(k):
                x[at:at + n] = blockset[i]
                at += n
        return x

    def export(self, outdir=os.getcwd(), prefix="sample", postfix=".txt", blocksize=float("inf"), delimiter="\t"):
        _export_blocks(self.flat(), outdir, prefix, postfix, blocksize, dict(delimiter=delimiter))  # :173-193

    def load(self, indir='', loadFile=None, prefix=None, postfix='.txt', nFiles=None, offset=1, delimiter='\t'):
        files = _file_list("sample", "a sample", indir, loadFile, prefix, postfix, nFiles, offset)  # :197-210
        x = numpy.vstack([numpy.loadtxt(open(f, "rb"), delimiter=delimiter) for f in files])  # :212-221
        n, k = self.n, self.k
        if x.shape == (2 * n, k):                # :225-227
            if not self.scaling:
                raise Exception("Loading a pre-generated, unscaled sample space requires that a 'scaling' function be defined.")
        elif x.shape == (2 * n * (1 + k), k):    # :229-246
            self.M_1 = x[0:n, ...]
            self.M_2 = x[n:2 * n, ...]
            self.N_j = x[2 * n:2 * n + k * n].reshape(k, n, k).copy()
            self.N_nj = x[2 * n + k * n:].reshape(k, n, k).copy()
        else:                                    # :247-248
            raise Exception("Loaded sample has shape " + str(x.shape) + ". Must have shape (%d,%d) or (%d,%d)."
                            % (2 * n, k, 2 * n * (1 + k), k))
        return x


def _file_list(noun, what, indir, loadFile, prefix, postfix, nFiles, offset):
    if loadFile:
        files = [os.path.join(indir, loadFile)]
    else:
        if not prefix:
            raise Exception("Either 'loadFile' or 'prefix' are required to load %s from file." % what)
        if not nFiles:
            raise Exception("Loading %s files with 'prefix' requires defining 'nFiles'." % noun)
        if prefix[-1] != "_":
            prefix += "_"
        files = [os.path.join(indir, prefix) + str(i) + postfix for i in range(offset, offset + nFiles)]
    for f in files:
        if not os.path.isfile(f):
            raise Exception("Cannot find input file " + f)
    return files


def _export_blocks(f, outdir, prefix, postfix, blocksize, savetxt_kw):
    # shared body of Sample.export (:173-193) and Objective.export (:393-413)
    blocksize = len(f) if blocksize > len(f) else int(blocksize)
    prefix = "_".join(str(prefix).split())
    if prefix[-1] == "_":
        prefix = prefix[:-1]
    prefix = os.path.join(outdir, prefix)
    nFiles = int(numpy.ceil(float(len(f)) / blocksize))
    if nFiles == 1:
        numpy.savetxt("%s%s" % (prefix, postfix), f, **savetxt_kw)
    else:
        for b in range(nFiles):
            numpy.savetxt("%s_%d%s" % (prefix, b + 1, postfix), f[b * blocksize:(b + 1) * blocksize], **savetxt_kw)


class Objective(object):
    def __init__(self, k, n, sample=None, objective_func=None, objective_vals=[], verbose=True, **loadArgs):
        self.k, self.n, self.sample = k, n, sample
        self.objective_func, self.verbose = objective_func, verbose
        if len(objective_vals) > 0:              # :297-298
            self.load(objective_vals)
        elif loadArgs:                           # :299-300
            self.load(**loadArgs)
        else:
            if not self.sample:                  # :302-305
                raise Exception("Generating a fresh objective requires that a 'sample' be defined.")
            elif not self.objective_func:
                raise Exception("Generating a fresh objective requires that an 'objective_func' be defined.")
            f = self.objective_func
            test = f(sample.M_1[0])              # :308 probe decides l
            try:
                l = len(test)
            except TypeError:
                l = 1
            self.fM_1 = numpy.zeros((n, l))      # :311-321
            self.fM_2 = numpy.zeros((n, l))
            self.fN_j = numpy.zeros((k, n, l))
            self.fN_nj = numpy.zeros((k, n, l))
            self.fM_1[0] = test                  # :329
            for i in range(1, n):                # :330-331
                self.fM_1[i] = f(sample.M_1[i])
            for i in range(n):                   # :336-337
                self.fM_2[i] = f(sample.M_2[i])
            for i in range(k):                   # :342-345
                for j in range(n):
                    self.fN_j[i][j] = f(sample.N_j[i][j])
            for i in range(k):                   # :350-353
                for j in range(n):
                    self.fN_nj[i][j] = f(sample.N_nj[i][j])

    def flat(self):
        # :357-391 -- same row order as Sample.flat
        n, k = self.n, self.k
        parts = [self.fM_1, self.fM_2] + [self.fN_j[i] for i in range(k)] + [self.fN_nj[i] for i in range(k)]
        return numpy.concatenate(parts, axis=0)

    def export(self, outdir=os.getcwd(), prefix="objective", postfix=".txt", blocksize=float("inf")):
        _export_blocks(self.flat(), outdir, prefix, postfix, blocksize, {})

    def load(self, obj_vals=[], indir='', loadFile=None, prefix=None, postfix='.txt', nFiles=None, offset=1, scaling=1.0):
        n, k = self.n, self.k
        if len(obj_vals) > 0:                    # :417-418
            x = obj_vals
        else:                                    # :420-447
            files = _file_list("objective", "an objective", indir, loadFile, prefix, postfix, nFiles, offset)
            obj = [numpy.loadtxt(open(f, "rb")) for f in files]
            one_observable = all(o.ndim == 1 for o in obj)   # :443 "list of 1-D arrays"
            x = numpy.hstack(obj) if one_observable else numpy.vstack(obj)
        if len(x) != 2 * n * (1 + k):            # :449, :471-472
            raise Exception("Loaded objective has length " + str(len(x)) + ". Must have length %d." % (2 * n * (1 + k)))
        x = numpy.asarray(x)
        if x.ndim == 1:                          # intended: one observable == (N,1); HEAD's :464 cannot
            x = x.reshape(-1, 1)                 # broadcast a 1-D slice into its (n,1) slot
        l = x.shape[1]                           # :455-461
        self.fM_1 = x[0:n, ...] / scaling        # :451
        self.fM_2 = x[n:2 * n, ...] / scaling
        self.fN_j = numpy.zeros((k, n, l))
        self.fN_nj = numpy.zeros((k, n, l))
        at = 2 * n
        for dst in (self.fN_j, self.fN_nj):      # :463-470
            for i in range(k):
                dst[i] = (x[at:at + n, ...] / scaling).reshape(n, l)
                at += n
        # :474-495 -- drop a row from all four matrices if any of them is NaN there
        isnan = numpy.logical_or(numpy.isnan(self.fM_1), numpy.isnan(self.fM_2))
        for i in range(k):
            isnan = numpy.logical_or(isnan, numpy.isnan(self.fN_j[i]))
            isnan = numpy.logical_or(isnan, numpy.isnan(self.fN_nj[i]))
        if isnan.ndim > 1:
            isnan = isnan[:, 0]                  # :481 first column only
        nans = [i for i in range(len(isnan)) if isnan[i]]
        self.fM_1 = numpy.delete(self.fM_1, nans, axis=0)
        self.fM_2 = numpy.delete(self.fM_2, nans, axis=0)
        self.fN_j = numpy.delete(self.fN_j, nans, axis=1)
        self.fN_nj = numpy.delete(self.fN_nj, nans, axis=1)
        if len(nans) > 0:
            print("WARNING: %d of %d objectives were NaN, %f%% loss" %
                  (len(nans), 2 * n * (1 + k), 100.0 * len(nans) / (2 * n * (1 + k))))


class Varsens(object):
    def __init__(self, objective, scaling_func=None, k=None, n=None, sample=None, verbose=True, literal_head=False):
        self.verbose = verbose
        if isinstance(sample, Sample):           # :548-551
            self.sample, self.k, self.n = sample, sample.k, sample.n
        elif k is not None and n is not None and scaling_func is not None:   # :552-555
            self.k, self.n = k, n
            if literal_head:
                self.sample = Sample(k, n, scaling_func, verbose)            # HEAD: verbose -> discard
            else:
                self.sample = Sample(k, n, scaling_func, verbose=verbose)
        elif not isinstance(objective, Objective):                           # :556-559
            raise ValueError("Must specify sample, (k,n,scaling_func), or Objective object")
        if isinstance(objective, Objective):     # :562-565
            self.objective, self.k, self.n = objective, objective.k, objective.n
        else:                                    # :567 (intended keyword form)
            self.objective = Objective(self.k, self.n, self.sample, objective, verbose=verbose)
        self.compute_varsens()

    def compute_varsens(self):
        o, n, k = self.objective, self.n, self.k
        self.E_2 = sum(o.fM_1 * o.fM_2) / n                                  # :577 sequential row sum
        self.var_y = numpy.var(numpy.concatenate((o.fM_1, o.fM_2), axis=0), axis=0, ddof=1)   # :583
        self.U_j = numpy.sum(o.fM_1 * o.fN_j, axis=1) / (n - 1)              # :591-593
        self.U_j += numpy.sum(o.fM_2 * o.fN_nj, axis=1) / (n - 1)
        self.U_j /= 2.0
        self.U_nj = numpy.sum(o.fM_1 * o.fN_nj, axis=1) / (n - 1)            # :594-596
        self.U_nj += numpy.sum(o.fM_2 * o.fN_j, axis=1) / (n - 1)
        self.U_nj /= 2.0
        shape = [k] if self.U_j.ndim == 1 else [k, self.U_j.shape[1]]        # :599-604
        self.sens, self.sens_t = numpy.zeros(shape), numpy.zeros(shape)
        for j in range(k):                                                   # :607-609
            self.sens[j] = (self.U_j[j] - self.E_2) / self.var_y
            self.sens_t[j] = 1.0 - ((self.U_nj[j] - self.E_2) / self.var_y)
        self.sens_2 = numpy.tensordot(o.fN_nj, o.fN_j, axes=([1], [1]))      # :612-616
        self.sens_2 += numpy.tensordot(o.fN_j, o.fN_nj, axes=([1], [1]))
        self.sens_2 /= 2.0 * (n - 1)
        self.sens_2 -= self.E_2
        self.sens_2 /= self.var_y
        self.sens_2n = numpy.tensordot(o.fN_nj, o.fN_nj, axes=([1], [1]))    # :618-622
        self.sens_2n += numpy.tensordot(o.fN_j, o.fN_j, axes=([1], [1]))
        self.sens_2n /= 2.0 * (n - 1)
        self.sens_2n -= self.E_2
        self.sens_2n /= self.var_y
