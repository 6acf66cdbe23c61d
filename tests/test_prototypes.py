"""The lane / register maps planned for the next transform kernel (tools/prototypes) stay correct."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pair_fft1024_lane_map():
    spec = importlib.util.spec_from_file_location(
        'pair_fft1024', os.path.join(ROOT, 'tools', 'prototypes', 'pair_fft1024.py'))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    assert module.main() < 1e-12
