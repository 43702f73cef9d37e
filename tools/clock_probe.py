"""Samples SM clock / power every 50 ms while the config-2 training step runs for a few seconds (is the step
power-capped?)."""
import os, subprocess, sys, time, statistics
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from magma_b200.config import MultimodalConfig
from magma_b200.magma import Magma
from magma_b200.train_loop import B200Engine
dev = torch.device("cuda:0")
mc = MultimodalConfig(batch_size=8, train_steps=1, encoder_name="clip_vit_large", adapter_config={"mlp": {"adapter_type": "normal", "downsample_factor": 4}}, image_seq_len=2, image_embed_dropout_prob=0.1, use_image_embed_layernorm=True, image_size=224, seq_len=128)
model = Magma(mc, device=dev); model.train(); eng = B200Engine(model, mc)
images = torch.randn(8, 3, 224, 224, device=dev).to(torch.bfloat16); captions = torch.randint(0, 50256, (8, 128), device=dev)
def step():
    out = eng(images, captions); eng.backward(out.loss); eng.step()
for _ in range(5): step()
torch.cuda.synchronize()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,temperature.gpu", "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"], stdout=subprocess.PIPE, text=True)
t0 = time.time(); n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 4.0:
    step(); n += 1
    if n % 8 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
p.terminate(); out = p.communicate()[0]
rows = [r.split(",") for r in out.strip().splitlines() if r.strip()]
clk = [float(r[0]) for r in rows]; pw = [float(r[1]) for r in rows]
print(f"steps {n}, ms/step {e0.elapsed_time(e1)/n:.2f}; samples {len(rows)}; sm clock median {statistics.median(clk)} min {min(clk)} max {max(clk)}; power median {statistics.median(pw)} max {max(pw)}; power_cap active in {sum('Active' in r[2] and 'Not' not in r[2] for r in rows)} samples; temp {rows[-1][4]}")
print("first 10:", rows[:10])
