"""GPU edge cases of the path: empty and tiny inputs, ragged tails (n % 4, n % 2048), zero / huge counts, all lines
discarded, chromosomes missing from the bias file, N >= 2^31 refused."""
import numpy as np
import pytest

from fithic_b200 import _capi, synth
from fithic_b200.engine import Biases, Contacts, Engine, Settings
from oracle import fithic_oracle as O
from tests.test_gpu_pipeline import run_engine
from tests.util import compare_pass, oracle_inputs

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("pval_impl")]


def both(contacts, frags, biases, st):
    got = run_engine(contacts, frags, biases, st)
    oc, fchr, fmid, fh, ost, ob = oracle_inputs(contacts, frags, st, biases)
    want = O.run_pipeline(oc, fchr, fmid, fh, ost, ob)
    assert len(got) == len(want)
    for r, o in zip(got, want):
        compare_pass(r, o)
    return got, want


@pytest.mark.parametrize("n", [2047, 2048, 2049, 4095, 6145, 10001])
def test_ragged_sizes(lib, n):
    contacts, frags, biases, _ = synth.make_intra(n, 100000, seed=n, chroms=["chr21", "chr22"], mean_count=6.0,
                                                  with_bias=True, inter_fraction=0.2)
    both(contacts, frags, biases, Settings(resolution=100000, noOfBins=20, noOfPasses=2))


def test_zero_and_huge_counts(lib):
    contacts, frags, biases, _ = synth.make_intra(50_000, 100000, seed=3, chroms=["chr20", "chr21"], mean_count=5.0)
    rng = np.random.default_rng(0)
    contacts.cnt[rng.integers(0, len(contacts), 500)] = 0          # count 0 -> k < 0 -> p = 1, distance still "seen"
    contacts.cnt[rng.integers(0, len(contacts), 20)] = 250_000     # beyond the shared-memory histogram path
    contacts.cnt[rng.integers(0, len(contacts), 3)] = 5_000_000    # beyond the lbeta table cap (per-contact fallback)
    both(contacts, frags, biases, Settings(resolution=100000, noOfBins=30))


def test_bias_file_without_a_chromosome_and_all_discarded(lib):
    contacts, frags, biases, _ = synth.make_intra(30_000, 100000, seed=5, chroms=["chr20", "chr21", "chr22"],
                                                  mean_count=5.0, with_bias=True, inter_fraction=0.2)
    # drop chr22 (id 2) from the bias table: every locus on it is "missing" -> -1 (fithic/fithic.py:1026-1031)
    off = biases.chr_off.copy()
    off[3] = off[2]
    b2 = Biases(biases.values[:off[2]].copy(), biases.mids[:off[2]].copy(), off)
    both(contacts, frags, b2, Settings(resolution=100000, noOfBins=30, allReg=True))
    # every bias discarded: every intra line gets p = 1, N stays what it is
    b3 = Biases(np.full_like(biases.values, -1.0), biases.mids, biases.chr_off)
    got, _ = both(contacts, frags, b3, Settings(resolution=100000, noOfBins=30))
    assert np.all(got[0]["p"] == 1.0) and np.all(got[0]["q"] == 1.0)


def test_single_chromosome_two_lines(lib):
    frags = synth.fragments_for(["chrA"], np.array([1_000_000]), 100000)
    c = Contacts(np.array([50000, 50000, 150000, 250000, 50000], np.int32), np.array([150000, 250000, 350000, 950000, 50000], np.int32),
                 np.array([5, 3, 2, 1, 9], np.int32), np.zeros(5, np.uint32), ["chrA"])
    both(c, frags, None, Settings(resolution=100000, noOfBins=4))


def test_empty_input_and_oversized_total(lib):
    frags = synth.fragments_for(["chrA"], np.array([1_000_000]), 100000)
    st = Settings(resolution=100000, noOfBins=4)
    eng = Engine(st, frags, None)
    eng.upload_contacts(Contacts(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.uint32),
                                 ["chrA"]))
    hist, present, scal = eng.hist_distance()
    torch.cuda.synchronize()
    assert int(hist.sum().item()) == 0 and int(scal.sum().item()) == 0
    q = eng.bh_qvalues(torch.empty(0, dtype=torch.float64, device="cuda"), 10.0)
    assert q.numel() == 0
    # sum of counts >= 2^31: scipy.special.bdtrc wraps n to int32 and the reference returns NaN / garbage
    c = Contacts(np.array([50000, 50000, 150000, 250000, 50000], np.int32),
                 np.array([150000, 250000, 350000, 950000, 50000], np.int32),
                 (np.array([5, 3, 2, 1, 9], np.int64) * 200_000_000).astype(np.int32), np.zeros(5, np.uint32), ["chrA"])
    eng.upload_contacts(c)
    with pytest.raises(_capi.FithicB200Error) as e:
        eng.run_pass(1, *eng.new_outlier_state())
    assert e.value.code == _capi.FHC_E_RANGE


def test_off_grid_distances_are_refused(lib):
    frags = synth.fragments_for(["chrA"], np.array([1_000_000]), 100000)
    c = Contacts(np.array([50000, 50000], np.int32), np.array([150000, 250001], np.int32), np.array([2, 3], np.int32),
                 np.zeros(2, np.uint32), ["chrA"])
    eng = Engine(Settings(resolution=100000, noOfBins=4), frags, None)
    eng.upload_contacts(c)
    with pytest.raises(ValueError, match="not a multiple of the resolution"):
        eng.run_pass(1, *eng.new_outlier_state())
