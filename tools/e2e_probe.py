#!/usr/bin/env python
"""Where does the end-to-end (host buffers) step lose time against the device-resident step?  Variants of the bench's
e2e loop built from the device API + explicit copies.  Diagnostic only."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multimodalvc_b200 import AVHubertConfig, AVHubertModel
from multimodalvc_b200 import build as avh_build


def main():
    avh_build.build()
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    model = AVHubertModel(AVHubertConfig.named("large")).to(dev, torch.bfloat16).eval()
    model.remove_pretraining_modules()
    B, T, S, NR, steps = 16, 150, 3, 4, 60
    hv = [torch.randn(B, 1, T, 88, 88).bfloat16().pin_memory() for _ in range(NR)]
    ha = [torch.randn(B, 104, T).bfloat16().pin_memory() for _ in range(NR)]
    dv = [[torch.empty_like(hv[0], device=dev) for _ in range(2)] for _ in range(S)]
    da = [[torch.empty_like(ha[0], device=dev) for _ in range(2)] for _ in range(S)]
    rv = [x.to(dev) for x in hv]
    ra = [x.to(dev) for x in ha]
    ho = [torch.empty(B, T, 1024, dtype=torch.bfloat16).pin_memory() for _ in range(S)]
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    copy_stream = torch.cuda.Stream(dev)
    main_s = torch.cuda.current_stream(dev)

    def run(name, h2d, d2h, host_sync, sep_copy=False, api_host=False):
        def step(i):
            st = streams[i % S]
            if host_sync:
                st.synchronize()
            if api_host:
                with torch.cuda.stream(st):
                    model.extract_finetune_host(hv[i % NR], ha[i % NR], None, out=ho[i % S], wait=False)
                return
            k = (i // S) & 1
            if h2d:
                v, a = dv[i % S][k], da[i % S][k]
                if sep_copy:
                    with torch.cuda.stream(copy_stream):
                        v.copy_(hv[i % NR], non_blocking=True)
                        a.copy_(ha[i % NR], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(copy_stream)
                    st.wait_event(ev)
                else:
                    with torch.cuda.stream(st):
                        v.copy_(hv[i % NR], non_blocking=True)
                        a.copy_(ha[i % NR], non_blocking=True)
            else:
                v, a = rv[i % NR], ra[i % NR]
            with torch.cuda.stream(st):
                y = model.extract_finetune({"audio": a, "video": v}, None)[0]
                if d2h:
                    ho[i % S].copy_(y, non_blocking=True)

        for i in range(2 * S):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in streams:
            st.wait_event(e0)
        t0 = time.time()
        for i in range(steps):
            step(i)
        t_enq = time.time() - t0
        for st in streams:
            main_s.wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(f"{name:46s} {ms:7.3f} ms/step  {B / ms * 1e3:7.1f} clips/s   host enqueue {t_enq / steps * 1e3:.3f} ms/step")

    run("device-resident, no host sync", False, False, False)
    run("device-resident, host sync per step", False, False, True)
    run("device inputs + D2H + host sync", False, True, True)
    run("H2D same stream + D2H + host sync", True, True, True)
    run("H2D same stream, no D2H, host sync", True, False, True)
    run("H2D on a copy stream + D2H + host sync", True, True, True, sep_copy=True)
    run("avh_forward_host_async (bench e2e)", True, True, True, api_host=True)


if __name__ == "__main__":
    main()
