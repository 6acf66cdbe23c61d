"""CPU checks of the restated mel filterbank (padertorch_b200/features.py: get_fbanks is paderbox's and not on
disk -- parity unpinned; these pin what can be pinned without it)."""
import numpy as np

from padertorch_b200 import features


def test_filterbank_shape_and_partition_of_unity():
    for htk in (False, True):
        fb = features.get_fbanks(16000, 1024, 80, 80, 7600, htk_mel=htk)
        assert fb.shape == (80, 513) and fb.dtype == np.float32
        assert fb.min() >= 0 and fb.max() <= 1.0 + 1e-6
        centres = features.mel_centre_frequencies(80, 80, 7600, htk)
        assert np.all(np.diff(centres) > 0) and abs(centres[0] - 80) < 1e-6 and abs(centres[-1] - 7600) < 1e-6
        freqs = np.arange(513) * 16000 / 1024
        inside = (freqs >= centres[1]) & (freqs <= centres[-2])
        # neighbouring triangles sum to one between the first and the last peak
        np.testing.assert_allclose(fb[:, inside].sum(0), 1.0, atol=1e-5)
        assert np.all(fb[:, freqs < centres[0]] == 0) and np.all(fb[:, freqs > centres[-1]] == 0)
        # every filter has a contiguous support around its own centre
        for i in range(80):
            support = np.flatnonzero(fb[i])
            if len(support):
                assert np.all(np.diff(support) == 1)
                assert centres[i] <= freqs[support[0]] and freqs[support[-1]] <= centres[i + 2]


def test_mel_scale_round_trip_and_anchors():
    f = np.array([0., 200., 1000., 4000., 7600.])
    for htk in (False, True):
        np.testing.assert_allclose(features.mel2hz(features.hz2mel(f, htk), htk), f, rtol=1e-10, atol=1e-9)
    assert abs(features.hz2mel(1000., htk_mel=True) - 1000.0) < 0.05       # HTK: 1000 Hz ~ 1000 mel
    assert abs(features.hz2mel(1000., htk_mel=False) - 15.0) < 1e-9        # Slaney: 1000 Hz = 15 mel


def test_slaney_normalisation_equalises_filter_areas():
    fb = features.get_fbanks(16000, 1024, 40, 0, 8000, htk_mel=False)
    normed = features.slaney_normalize(fb, 16000, 1024, 0, 8000)
    areas = normed.sum(1) * (16000 / 1024)      # integral over frequency
    np.testing.assert_allclose(areas[3:-1], 1.0, rtol=0.05)
