'''Helpers shared by the tests: load golden fixtures, compare CSR triples.'''
import os

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
GEN = os.path.join(HERE, 'golden', 'generated')
REF = os.path.join(HERE, 'golden', 'ref_data')


def load_case(name):
    return numpy.load(os.path.join(GEN, name + '.npz'))


def read_ref_matrix(fname, n):
    '''Reader for the reference's 1-based `row col value` text format (tests/test_fvm.py:386-407).'''
    rows, cols, vals = [], [], []
    with open(os.path.join(REF, fname)) as f:
        for line in f:
            parts = line.split()
            if not parts:
                continue
            rows.append(int(parts[0]) - 1)
            cols.append(int(parts[1]) - 1)
            vals.append(float(parts[2]))
    rows = numpy.array(rows)
    assert numpy.all(numpy.diff(rows) >= 0)
    begA = numpy.zeros(n + 1, dtype=numpy.int64)
    numpy.add.at(begA, rows + 1, 1)
    return numpy.array(vals), numpy.array(cols, dtype=numpy.int64), numpy.cumsum(begA)


def read_ref_vector(fname):
    with open(os.path.join(REF, fname)) as f:
        return numpy.array([float(line) for line in f if line.strip()])


def compress(vals, cols, row_ptr, tol=1e-14):
    '''Drop |v| <= 1e-14 exactly like CrsMatrix.compress (CrsMatrix.py:65) so that a fixed
    structural pattern with explicit zeros can be compared with the reference's
    value-dependent pattern.  Host-side test helper.'''
    keep = numpy.abs(vals) > tol
    counts = numpy.add.reduceat(keep, row_ptr[:-1]) if len(vals) else numpy.zeros(len(row_ptr) - 1, int)
    counts = numpy.where(numpy.diff(row_ptr) > 0, counts, 0)
    new_ptr = numpy.zeros(len(row_ptr), dtype=numpy.int64)
    new_ptr[1:] = numpy.cumsum(counts)
    return vals[keep], cols[keep], new_ptr


def assert_csr_equal(got, want, rtol=0.0, what=''):
    gv, gc, gp = got
    wv, wc, wp = want
    assert numpy.array_equal(numpy.asarray(gp, dtype=numpy.int64), numpy.asarray(wp, dtype=numpy.int64)), what + ' row pointers differ'
    assert numpy.array_equal(numpy.asarray(gc, dtype=numpy.int64), numpy.asarray(wc, dtype=numpy.int64)), what + ' column indices differ'
    if rtol == 0.0:
        assert numpy.array_equal(gv, wv), what + ' values not bit-identical (max rel %.3e)' % (
            numpy.max(numpy.abs(gv - wv) / numpy.abs(wv)) if len(wv) else 0)
    else:
        assert numpy.all(numpy.abs(gv - wv) <= rtol * numpy.abs(wv)), what + ' values differ (max rel %.3e)' % (
            numpy.max(numpy.abs(gv - wv) / numpy.abs(wv)))
