"""Host-side pieces of bench.py that need no GPU: the clock sampler's sample selection."""
import datetime
import time

import bench


def test_clock_sampler_uses_the_samples_inside_the_timed_region():
    s = bench.ClockSampler(0)
    s.proc = object.__new__(type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0, "kill": lambda self: None}))
    now = time.time()

    def line(t, mhz, slow="Not Active"):
        stamp = datetime.datetime.fromtimestamp(t).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
        return "%s, 0, %d, 1965, 400.0, 0x0, %s, Not Active, Not Active, Not Active" % (stamp, mhz, slow)

    assert abs(bench.ClockSampler._stamp(line(now, 1).split(",")[0]) - now) < 2e-3
    s.lines = [line(now - 1.0, 345), line(now - 0.5, 345, "Active"),          # idle samples before the region (one with a throttle flag)
               line(now + 0.05, 1965), line(now + 0.10, 1950), line(now + 0.15, 1965),
               line(now + 1.0, 345), "garbage"]
    s.t0, s.t1 = now, now + 0.2
    out = s.stop()
    assert out["samples_inside_region"] == 3 and out["samples"] == 3
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == []
    # a region shorter than the sampling period: the samples next to it are reported instead of none
    s.t0, s.t1 = now + 0.51, now + 0.52
    out = s.stop()
    assert out["samples_inside_region"] == 0 and out["samples"] == 3
