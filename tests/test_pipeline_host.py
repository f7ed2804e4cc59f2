"""Host logic of the upload / run / fetch software pipeline (PipelinedBatches) with a stand-in for the
device handle: ordering, exclusive use of every handle, error propagation without deadlock."""
import random
import threading
import time

import numpy as np
import pytest

import soundgen_beta_b200.api as api


class FakeBatch:
    fail_on_run = -1
    runs = 0
    lock = threading.Lock()

    def __init__(self):
        self.desc = None
        self.busy = threading.Lock()
        self.log = []

    def _use(self, what):
        assert self.busy.acquire(blocking=False), 'a handle was used by two stages at once'
        time.sleep(random.uniform(0, 0.002))
        self.log.append(what)
        self.busy.release()

    def upload(self, d):
        self._use('u')
        self.desc = d

    def run(self):
        with FakeBatch.lock:
            FakeBatch.runs += 1
            n = FakeBatch.runs
        if n == FakeBatch.fail_on_run:
            raise api.SoundgenError(-2, 'injected failure')
        self._use('r')

    def lengths(self):
        return np.array([10])

    def fetch(self, dtype, out=None):
        self._use('f')
        return [out]

    def close(self):
        pass


class FakeLib:
    def sgb_pin(self, *a):
        return 0

    def sgb_unpin(self, *a):
        return 0


@pytest.fixture
def fake(monkeypatch):
    monkeypatch.setattr(api, 'Batch', FakeBatch)
    monkeypatch.setattr(api._abi, 'load', lambda: FakeLib())
    FakeBatch.fail_on_run, FakeBatch.runs = -1, 0
    return FakeBatch


@pytest.mark.parametrize('npipe,runners', [(8, 1), (8, 3), (3, 4), (1, 2), (16, 2)])
def test_every_sub_batch_goes_through_the_three_stages_in_order(fake, npipe, runners):
    p = api.PipelinedBatches(list(range(npipe)), runners=runners)
    p.run_steps(3)
    for b in p.batches:
        assert ''.join(b.log) == 'urf' * 3
    p.run_steps(2, transfer=False)          # resident: no upload, no fetch
    for b in p.batches:
        assert ''.join(b.log) == 'urf' * 3 + 'rr'
    assert len(p.step()) == npipe
    p.close()


def test_a_failing_stage_surfaces_and_nothing_hangs(fake):
    p = api.PipelinedBatches(list(range(8)), runners=3)
    fake.fail_on_run = 11
    t = time.time()
    with pytest.raises(api.SoundgenError, match='injected failure'):
        p.run_steps(4)
    assert time.time() - t < 10
    fake.fail_on_run = -1
    p.run_steps(1)                          # the pipeline is usable again
