"""Timeline of one msda_forward_backward_host call on config 2 (PAVENET_MSDA_TRACE_E2E): when each
upload, kernel pair and download finished, relative to the start of the call.
Usage: python tools/trace_e2e.py [piece_mib]"""
import os
import sys
import time

os.environ['PAVENET_MSDA_TRACE_E2E'] = '/tmp/e2e_trace.csv'
import torch  # noqa: E402

sys.path.insert(0, '.')
import bench  # noqa: E402
import pavenet_b200  # noqa: E402


def main():
    piece = float(sys.argv[1]) if len(sys.argv) > 1 else 0
    p = bench.make_problem('encoder_cfg2', seed=0, device=torch.device('cuda', 0))
    host = {k: torch.empty(p[k].shape, dtype=p[k].dtype).pin_memory() for k in ('value', 'loc', 'aw', 'grad_out')}
    for k in host:
        host[k].copy_(p[k])
    d = p['dims']
    out_h = torch.empty((d['B'], d['Q'], d['M'] * d['D'])).pin_memory()
    gv_h = torch.empty(p['value'].shape).pin_memory()
    gl_h = torch.empty(p['loc'].shape).pin_memory()
    ga_h = torch.empty(p['aw'].shape).pin_memory()
    shapes_h, lsi_h = p['shapes'].cpu(), p['lsi'].cpu()
    hws = pavenet_b200.HostWorkspace()
    if piece > 0:
        hws.set_piece_bytes(int(piece * (1 << 20)))

    def call():
        hws.forward_backward(host['value'], shapes_h, lsi_h, host['loc'], host['aw'], host['grad_out'],
                             out=out_h, grad_value=gv_h, grad_sampling_loc=gl_h, grad_attn_weight=ga_h)
    for _ in range(3):
        call()
    t0 = time.perf_counter()
    for _ in range(10):
        call()
    print('piece %s MiB: %.3f ms per call (wall, 10 calls, tracing on)' % (piece or 'default', (time.perf_counter() - t0) * 100))
    rows = [l.strip().split(',') for l in open('/tmp/e2e_trace.csv')][1:]
    last = {}
    for kind, b, pc, ms in rows:
        last[kind] = float(ms)
    print('stage ends (ms): ' + '  '.join('%s %.3f' % (k, v) for k, v in last.items()))
    # compact timeline: per piece, when inputs arrived / compute done / outputs down
    per = {}
    for kind, b, pc, ms in rows:
        per.setdefault((int(b), int(pc)), {})[kind] = float(ms)
    for (b, pc), m in sorted(per.items()):
        print('b%d piece %2d  ' % (b, pc) + '  '.join('%s %6.3f' % (k, m[k]) for k in 'VICOG' if k in m))


if __name__ == '__main__':
    main()
