"""Where does the gradient error of the fused / tape GPU paths sit relative to an fp64 oracle?
(The fp32 CPU oracle has rounding noise of its own.)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import truth_recommendation_gnn_b200 as trg
from truth_recommendation_gnn_b200 import synth, fused_step
from oracle import sage as osage
from tests.util import oracle_model

dev = torch.device("cuda:0")
verbose = len(sys.argv) > 1
for layers, h in ((1, 128), (2, 128), (2, 64), (3, 32), (1, 64)):
    U, P, Ee, Es = 400, 900, 9000, 2500
    g = synth.synth_graph(U, P, Ee, Es, h, seed=11, skew=True)
    sd = synth.init_state_dict(h, h, layers)
    neg = synth.synth_neg(P, Ee, 2)
    grads, outs = {}, {}
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        ref = oracle_model(h, layers, sd).to(dt)
        out = ref({k: v.to(dt) for k, v in g.x_dict.items()}, g.edge_index_dict)
        l = osage.link_loss(out["user"], out["post"], g.train_edge_index[0], g.train_edge_index[1], neg,
                            g.interaction_type_tensor.to(dt), U)
        l.backward()
        grads[name] = {n: p.grad.double() for n, p in ref.named_parameters()}
        outs[name] = {k: v.detach().double() for k, v in out.items()}
    gd = g.to(dev)
    ma = trg.StackedWeightedRGCN(h, layers) if layers > 1 else trg.WeightedRGCN(h)
    ma.load_state_dict(sd); ma = ma.to(dev)
    with torch.no_grad():
        o = ma(gd.x_dict, gd.edge_index_dict)
    for k in ("user", "post"):
        t = outs["f64"][k]
        print(f"  L={layers} H={h} emb {k}: gpu {float((o[k].double().cpu()-t).abs().max()/t.abs().max()):.2e} "
              f"cpu32 {float((outs['f32'][k]-t).abs().max()/t.abs().max()):.2e}")
    fused_step.loss_and_grads(ma, gd.x_dict, gd.edge_index_dict, gd.train_edge_index, gd.interaction_type_tensor, U, neg.to(dev))
    worst = [0, 0, 0]
    for n, p in ma.named_parameters():
        t = grads["f64"][n]; s = t.abs().max()
        e_gpu = float((p.grad.double().cpu() - t).abs().max() / s)
        e_cpu = float((grads["f32"][n] - t).abs().max() / s)
        e_gc = float((p.grad.double().cpu() - grads["f32"][n]).abs().max() / s)
        if verbose:
            print(f"    {n}: gpu {e_gpu:.2e} cpu32 {e_cpu:.2e}  scale {float(s):.2e}")
        worst = [max(worst[0], e_gpu), max(worst[1], e_cpu), max(worst[2], e_gc)]
    print(f"L={layers} H={h}: worst rel err vs fp64: gpu {worst[0]:.2e}, cpu-f32 oracle {worst[1]:.2e}; gpu vs cpu-f32 {worst[2]:.2e}")
