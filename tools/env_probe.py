import torch, time, json, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_hanoi_b200.engine import VecHanoi
from muzero_hanoi_b200 import _lib
b=1<<24
for n in (10,5):
  env=VecHanoi(n,200,b); env.reset()
  def timeit(f, iters=20):
      for _ in range(3): f()
      torch.cuda.synchronize()
      e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
      e0.record()
      for i in range(iters): f()
      e1.record(); torch.cuda.synchronize()
      return e0.elapsed_time(e1)/iters
  acts=torch.randint(0,6,(b,),dtype=torch.uint8,device='cuda')
  t=timeit(lambda: env.step(acts,want_obs=False)); print('N=%d'%n, 'step 14B: %.1f us  %.3e steps/s  %.0f GB/s'%(t*1e3,b/t*1e3,14*b/t/1e6))
  k=[0]
  def sr():
      env.step_random(seed=1,step_index=k[0]); k[0]+=1
  t=timeit(sr); print('step_random 13B: %.1f us %.3e steps/s %.0f GB/s'%(t*1e3,b/t*1e3,13*b/t/1e6))
  for K in (16,64,256):
      t=timeit(lambda: env.rollout_random(K,seed=1,step_index=0),iters=5); print('rollout K=%d: %.1f us %.3e steps/s'%(K,t*1e3,b*K/t*1e3))
