"""Host-side (CPU enqueue) cost of one train step: cProfile over steps on a small graph, where the GPU
is faster than the launch path, so the time measured is launch/bookkeeping overhead only."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import truth_recommendation_gnn_b200 as trg
from truth_recommendation_gnn_b200 import synth, _lib

dev = torch.device("cuda:0")
U, P, Ee, Es, H, L = 10_000, 50_000, 400_000, 100_000, 128, 2
g = synth.synth_graph(U, P, Ee, Es, H, seed=0, device=dev)
model = trg.StackedWeightedRGCN(H, L); model.load_state_dict(synth.init_state_dict(H, H, L)); model = model.to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
negs = [synth.synth_neg(P, Ee, i, device=dev) for i in range(4)]
def step(i):
    return trg.train_step(model, opt, g.x_dict, g.edge_index_dict, g.train_edge_index, g.interaction_type_tensor,
                          U, P, neg_p=negs[i % 4], return_tensor=True)
for i in range(10): step(i)
torch.cuda.synchronize()
for prof_on in (False, True):
    _lib.PROF.enabled = prof_on; _lib.PROF.reset()
    t0 = time.perf_counter()
    for i in range(50): step(i)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"PROF={prof_on}: enqueue {1e3*(t1-t0)/50:.3f} ms/step, +sync {1e3*(t2-t0)/50:.3f} ms/step")
_lib.PROF.enabled = False
pr = cProfile.Profile(); pr.enable()
for i in range(50): step(i)
pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(22)
