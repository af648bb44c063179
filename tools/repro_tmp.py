import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np, torch
from kubernetes_acs_engine_autoscaler_b200 import synthetic as syn
from kubernetes_acs_engine_autoscaler_b200.engine import Engine
eng=Engine(0, watchdog_ms=5000)
c=syn.make_cluster(4097,1025,8,8,seed=4,free_frac=0.25)
used0=syn.initial_used(c)
d_used=eng.dev(used0,torch.float64)
placed,dec=eng.first_fit_nodes(eng.dev(c['req'],torch.float64),None,eng.dev(c['cap_type'],torch.float64),eng.dev(c['node_type'],torch.int32),d_used)
torch.cuda.synchronize(); print('ok', int(dec.item()))
