"""TensorBoard scalar summaries without TensorFlow (reference train.py:185-189,227-228,
260-262): ``FileWriter(logdir).add_scalars(step, {'summaries/loss': v, ...})`` writes a
standard ``events.out.tfevents.*`` file through the ``tensorboard`` package's record writer."""
from __future__ import annotations

import time


class FileWriter:
    def __init__(self, logdir: str):
        from tensorboard.summary.writer.event_file_writer import EventFileWriter
        self._w = EventFileWriter(logdir)

    def add_scalars(self, step: int, scalars: dict):
        from tensorboard.compat.proto.event_pb2 import Event
        from tensorboard.compat.proto.summary_pb2 import Summary
        s = Summary(value=[Summary.Value(tag=k, simple_value=float(v)) for k, v in scalars.items()])
        self._w.add_event(Event(wall_time=time.time(), step=int(step), summary=s))

    def flush(self):
        self._w.flush()

    def close(self):
        self._w.close()
