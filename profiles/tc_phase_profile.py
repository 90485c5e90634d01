import torch, ctypes as C, sys
sys.path.insert(0,'/root/repo')
import ddpg_trucktrailer_b200 as tt
from ddpg_trucktrailer_b200 import _lib
L = tt.load()
N = 1<<22
actor = tt.agent.CudaActor(); actor.load_state_dict(tt.init_actor_state_dict(seed=0))
obs = torch.empty(N,23,device='cuda').uniform_(-1,1)
dbg = torch.zeros(24, dtype=torch.int64, device='cuda')
L.tt_debug_set_tc_profile.argtypes=[C.c_void_p]
for prec in ("f16",):
    for _ in range(3): actor.forward(obs, precision=prec)
    L.tt_debug_set_tc_profile(dbg.data_ptr())
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); actor.forward(obs, precision=prec); e1.record(); torch.cuda.synchronize()
    d = dbg.cpu().tolist(); nt = d[3]
    print(prec, 'ms', e0.elapsed_time(e1), 'tiles/blk', nt)
    print('  layer-2 MMA thread waits per tile: h2free %.0f  a2full %.0f  w2full %.0f' % (d[0]/nt, d[1]/nt, d[2]/nt))
    print('  epilogue per tile (v4): wait_h2a %.0f  pass1a %.0f  epi1(+waits) %.0f  wait_h2b %.0f  epi2 %.0f  total %.0f  stage %.0f' % tuple(x/nt for x in d[4:11]))
    print('  epi2 detail: pass1b %.0f  exchange+reloadA %.0f  pass2A(+reload B) %.0f  pass2B %.0f' % tuple(x/nt for x in d[11:15]))
    print('  epilogue waits: layer-1 MMAs (WFULL) %.0f  A2 blocks (A2FREE) %.0f  statistics barrier %.0f  output barrier %.0f' % tuple(x/nt for x in d[15:19]))
    L.tt_debug_set_tc_profile(None)
