"""`import optax` (ofdft_normflows/utils.py:8) -- only the learning-rate schedules of get_scheduler (utils.py:147-184)
are touched, and only when that function is called (it is outside the hot path)."""


def constant_schedule(value):
    return lambda step: value


def _outside(*a, **k):
    raise NotImplementedError("optax is outside the hot path; only constant_schedule is restated")


warmup_cosine_decay_schedule = join_schedules = cosine_onecycle_schedule = _outside
