"""Diagnostic (test infrastructure, run by hand: python tests/grad_table.py [bf16|fp32]): per-parameter gradient error of the fused train step vs the oracle (same-sign backward)."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import synth
from oracle import larva_oracle as O

def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
    blocks = [2, 1]
    params = synth.make_larva_params(blocks, seed=11, bias_std=0.02)
    lr, hr = synth.make_images(2, 12, 10, seed=111)
    m = importlib.import_module('models.LarvaNet').create_model()
    m.parse_args(['--num_modules=2', '--num_blocks=2,1', f'--precision={prec}'])
    m.prepare(is_training=True, scales=[4])
    m.get_model().load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    eng = m._engine()
    loss = eng.train_step(torch.from_numpy(lr).cuda(), torch.from_numpy(hr).cuda(), keep_exits=True).item()
    ex = [e.cpu().numpy() for e in eng.last_exits]
    rl, rg, ro = O.larvanet_train_step(params, lr, hr, blocks, sign_from=ex, tapes_from=eng.saved_activations())
    print('loss', loss, rl)
    for n, p in m.get_model().named_parameters():
        g = p.grad.cpu().numpy().astype(np.float64)
        r = rg[n]
        print(f'{n:45s} rel={np.linalg.norm(g-r)/np.linalg.norm(r):.4f} |ref|={np.linalg.norm(r):.3e} |got|={np.linalg.norm(g):.3e} '
              f'cos={np.sum(g*r)/np.linalg.norm(g)/np.linalg.norm(r):.5f}')
    # head: compare against oracle head wgrad using the DEVICE dfin (isolates the head wgrad kernel)
    b = eng._train.entries[(2, 12, 10, True)]['bufs']
    # planar-8 [n, h, c/8, w, 8] -> NCHW
    dfin = b.dfin[0].float().cpu().permute(0, 2, 4, 1, 3).reshape(2, 48, 12, 10).numpy().astype(np.float64)
    _, dwh, dbh = O.conv2d_backward(lr.astype(np.float64), np.zeros((48, 3, 3, 3)), dfin)
    scale = b.scale
    gh = m.get_model().head.feature_extraction.weight.grad.cpu().numpy()
    print('head wgrad kernel vs oracle on device dfin: rel', np.linalg.norm(gh - dwh * scale) / np.linalg.norm(dwh * scale))

main()
