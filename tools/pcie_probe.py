"""Developer tool (GPU box): what the PCIe link gives for the e2e step's copies -- one 44.8 MB host->device copy vs the
five attribute arrays as separate copies, with and without the 24.9 MB image download running the other way."""
import json
import torch

dev = torch.device("cuda:0")
sizes = [9599484, 9599484, 12799312, 3199828, 9599484]  # means, scales, rotations, opacities, packed SH (bytes)
tot = sum(sizes)
h_one = torch.empty(tot, dtype=torch.uint8).pin_memory()
h_parts = [torch.empty(s, dtype=torch.uint8).pin_memory() for s in sizes]
d_one = torch.empty(tot, dtype=torch.uint8, device=dev)
d_parts = [torch.empty(s, dtype=torch.uint8, device=dev) for s in sizes]
img_d = torch.empty(24883200, dtype=torch.uint8, device=dev)
img_h = torch.empty(24883200, dtype=torch.uint8).pin_memory()
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()


def run(split, down, n=40):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        with torch.cuda.stream(s_up):
            if split:
                for d, h in zip(d_parts, h_parts):
                    d.copy_(h, non_blocking=True)
            else:
                d_one.copy_(h_one, non_blocking=True)
        if down:
            with torch.cuda.stream(s_dn):
                img_h.copy_(img_d, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s_up)
    torch.cuda.current_stream().wait_stream(s_dn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return {"ms_per_step": round(ms, 4), "steps_per_s": round(1e3 / ms, 1), "h2d_GBps": round(tot / ms / 1e6, 1)}


for split in (False, True):
    for down in (False, True):
        run(split, down, 5)
        print(json.dumps({"five_copies": split, "with_image_download": down, **run(split, down)}))
