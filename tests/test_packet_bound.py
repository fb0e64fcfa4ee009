"""CPU: the chord bound behind RTC_FLAG_PACKET (csrc/rtc_trace.cu: test_packet_groups, packet_eps), checked
numerically in binary64: whenever some ray of a packet reaches the per-ray filter's threshold, one of the packet's two END
rays reaches the threshold of the wider deflation E = 1e-5 + 1.5 D^2 -- and the check has teeth: with the D^2 term removed
it finds violations.  (The GPU tests compare the packet kernel's hits with the oracle's on whole frames.)"""
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    spec = importlib.util.spec_from_file_location("check_packet_bound", os.path.join(HERE, "..", "scripts", "experiments", "check_packet_bound.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_packet_chord_bound_is_conservative():
    m = load()
    bad = 0
    for args in ((2160, 3840, 0.325, 0.577), (150, 399, 0.866, 0.577), (4320, 7680, 0.325, 0.577)):
        for kr in (8, 4):
            bad += m.check(*args, n_pk=4000, n_sph=150, kr=kr, tiny=False)
            bad += m.check(*args, n_pk=4000, n_sph=150, kr=kr, tiny=True)
    assert bad == 0


def test_packet_chord_bound_check_has_teeth():
    m = load()
    assert m.check(2160, 3840, 0.325, 0.577, n_pk=4000, n_sph=1500, tiny=True, coeff=0.0, base=2.7e-6) > 0
