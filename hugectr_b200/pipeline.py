"""Pipeline scheduler: multi-stream DAG of schedulable stages with optional whole-pipeline CUDA-graph
capture.  API parity with HugeCTR/include/pipeline.hpp:28-108 (Scheduleable,
StreamContextScheduleable{set_stream, set_absolute_stream, wait_event, record_done},
GraphScheduleable, Pipeline{run, run_graph}); implementation on torch streams / events / CUDAGraph.
``Model.eval`` runs through it (embedding forward beside the bottom network, top network behind both, the
whole evaluation step captured as one CUDA graph); ``Model._step_body`` is the hand-scheduled training
instance of the same idea (it additionally interleaves the bucketed all-reduce hooks).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch


class Scheduleable:
    def run(self, ctx: "Pipeline"):
        raise NotImplementedError


class StreamContextScheduleable(Scheduleable):
    def __init__(self, fn: Callable[[], None], name: str = ""):
        self.fn, self.name = fn, name
        self.stream_name: Optional[str] = None
        self.absolute = False
        self.waits: List["StreamContextScheduleable"] = []
        self.done_event = None
        self.record = False

    def set_stream(self, name: str):
        self.stream_name = name
        return self

    def set_absolute_stream(self, name: str):
        """a stream that is not joined back into the capture (prefetch across iterations)"""
        self.stream_name, self.absolute = name, True
        return self

    def wait_event(self, others):
        self.waits.extend(others if isinstance(others, (list, tuple)) else [others])
        return self

    def record_done(self):
        self.record = True
        return self

    def run(self, ctx: "Pipeline"):
        cuda = torch.cuda.is_available() and ctx.device.type == "cuda"
        if not cuda:
            self.fn()
            return
        cur = torch.cuda.current_stream()
        stream = ctx.stream(self.stream_name) if self.stream_name else cur
        if stream is not cur and not self.absolute:
            stream.wait_stream(cur)          # a side stage starts after everything queued before the pipeline
        for w in self.waits:
            if w.done_event is not None:
                stream.wait_event(w.done_event)
        with torch.cuda.stream(stream):
            self.fn()
            if self.record or True:
                self.done_event = torch.cuda.Event()
                self.done_event.record(stream)


class GraphScheduleable(Scheduleable):
    """A list of stages captured and replayed as one CUDA graph."""

    def __init__(self, stages: List[Scheduleable]):
        self.stages = stages
        self.graph = None
        self.warm = 0

    def run(self, ctx: "Pipeline", use_graph: bool = True):
        cuda = torch.cuda.is_available() and ctx.device.type == "cuda"
        if not (cuda and use_graph):
            for s in self.stages:
                s.run(ctx)
            return
        if self.graph is None:
            if self.warm < 2:
                for s in self.stages:
                    s.run(ctx)
                self.warm += 1
                return
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                main = torch.cuda.current_stream()
                for s in self.stages:
                    s.run(ctx)
                for st in ctx.streams.values():
                    main.wait_stream(st)
            self.graph = g
        self.graph.replay()


class Pipeline:
    def __init__(self, name: str, device, stages: List[Scheduleable]):
        self.name, self.device, self.stages = name, torch.device(device), stages
        self.streams: Dict[str, "torch.cuda.Stream"] = {}
        self._graph = GraphScheduleable(stages)

    def stream(self, name: str):
        if name not in self.streams:
            self.streams[name] = torch.cuda.Stream(self.device)
            self.streams[name].wait_stream(torch.cuda.current_stream())
        return self.streams[name]

    def run(self):
        for s in self.stages:
            s.run(self)
        if self.device.type == "cuda":
            main = torch.cuda.current_stream()
            for st in self.streams.values():
                main.wait_stream(st)

    def run_graph(self):
        self._graph.run(self, True)
