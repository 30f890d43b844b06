"""Host-side profile of the end-to-end training step (plugin call with host batches): cProfile top entries + phase timers."""
import cProfile, importlib, os, pstats, sys, time, types
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import synth
from larvanet_b200.prefetch import DevicePrefetcher

dev = torch.device('cuda', 0)
model = importlib.import_module('models.LarvaNet').create_model()
model.parse_args(['--num_modules=4', '--num_blocks=4,4,4,4', '--precision=bf16'])
model.volume_per_step = 48 * 48 * 16 * 3
model.prepare(is_training=True, scales=[4])
model.args.val_volume = 1e30
model.global_step = 1
pool = [synth.make_images(16, 48, 48, seed=100 + i) for i in range(8)]
host = [(torch.from_numpy(l).pin_memory(), torch.from_numpy(h).pin_memory()) for l, h in pool]


def batches():
    i = 0
    while True:
        yield host[i % 8]
        i += 1


feeder = DevicePrefetcher(batches(), dev, depth=2, defer=True)
ns = types.SimpleNamespace(train_path='/tmp')


def step():
    x, t = next(feeder)
    return model.train_step_larva(ns, None, x, t)


for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    step()
torch.cuda.synchronize()
print(f'e2e step {(time.perf_counter() - t0) / 300 * 1e6:.1f} us')
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('cumulative').print_stats(28)

# phase timers without the profiler
import collections
acc = collections.Counter()
eng = model._engine()
for _ in range(300):
    t0 = time.perf_counter()
    x, t = next(feeder)
    t1 = time.perf_counter()
    loss = eng.train_step(x, t)
    t2 = time.perf_counter()
    model.optim.step()
    t3 = time.perf_counter()
    v = loss.item()
    t4 = time.perf_counter()
    acc['feeder'] += t1 - t0; acc['train_step enqueue'] += t2 - t1; acc['optim enqueue'] += t3 - t2; acc['item (wait + D2H)'] += t4 - t3
print({k: f'{v / 300 * 1e6:.1f} us' for k, v in acc.items()})
