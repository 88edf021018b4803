"""The two CPU evaluations of the VQT front end against each other and against the product's filter design (all on the CPU):

* oracle/vqt_oracle.py (time-domain definition) vs oracle/vqt_recursive_oracle.py (librosa's octave-recursive, sparsified computation):
  same transform at spectral peaks, systematically different in quiet low-octave bins -- the bound on how far the plain definition is
  from what librosa computes;
* piano_a2s_b200.vqt.design_filters_librosa (the recursion composed analytically into ONE direct-form filter bank, the product's
  filters) evaluated with numpy vs the recursive oracle, which really decimates and takes FFTs: 1e-6 in the log domain."""
import numpy as np

from oracle import vqt_oracle as VO
from oracle import vqt_recursive_oracle as VR


def _signal(n=24000, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    return (0.3 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 1318.5 * t * (1 + 0.01 * t)) + 0.1 * np.sin(2 * np.pi * 82.4 * t)
            + 0.02 * rng.standard_normal(t.size))


def test_direct_definition_vs_librosa_recursion():
    y = _signal()
    a, b = VO.vqt_magnitude(y), VR.vqt_magnitude(y)
    k = np.unravel_index(np.argmax(a), a.shape)
    assert abs(a[k] - b[k]) / a[k] < 5e-3                              # same transform, same scaling, at the spectral peak
    la, lb = VO.get_vqt(y, algorithm="direct"), VO.get_vqt(y, algorithm="librosa")
    d = np.abs(la - lb)
    loud = la >= 0.75                                                  # within 20 dB of the clip maximum
    print("loud bins: mean %.4f max %.4f; all bins: mean %.4f max %.4f" % (d[loud].mean(), d[loud].max(), d.mean(), d.max()))
    assert d[loud].mean() < 1e-2 and d.mean() < 2e-2
    assert d.max() > 0.1                                               # ... and NOT the same in quiet low-octave bins (side-lobe leakage)


def test_composed_filter_bank_equals_the_recursion():
    from piano_a2s_b200.vqt import design_filters_librosa
    W, p_min = design_filters_librosa()
    y = _signal(16000, seed=3)
    K, hop = W.shape[1], 160
    T = 1 + len(y) // hop
    ypad = np.concatenate([np.zeros(-p_min), y, np.zeros(K + hop)])
    C = ypad[np.arange(T)[:, None] * hop + np.arange(K)[None, :]] @ W.T.astype(np.float64)
    mag = np.sqrt(C[:, 0::2] ** 2 + C[:, 1::2] ** 2).T
    ref = VR.vqt_magnitude(y)                                          # extend=True: the infinite-signal convention of the product
    assert mag.shape == ref.shape
    assert np.abs(mag - ref).max() < 1e-6 * ref.max()
    lit = VR.vqt_magnitude(y, extend=False)                            # the literal, truncating recursion differs only at the clip edges
    assert np.abs(mag - lit)[:, 8:-8].max() < 1e-6 * ref.max()
    got = (VO.amplitude_to_db(mag) / 80.0 + 1.0).T
    assert np.abs(got - VR.get_vqt(y)).max() < 2e-5
