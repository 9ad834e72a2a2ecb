import importlib, os, sys, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/examples")
import torch
importlib.import_module("unimol")
from unicore import options, tasks, utils
from unicore.trainer import Trainer
flags = ["--task","synthetic_unimol","--loss","unimol","--arch","unimol_base","--synthetic-num-samples","256","--optimizer","adam","--adam-betas","(0.9, 0.99)","--clip-norm","1.0","--lr","1e-4","--lr-scheduler","polynomial_decay","--warmup-updates","100","--total-num-update","100000","--max-update","100000","--batch-size","32","--seed","1","--no-save","--disable-validation","--log-format","none","--distributed-world-size","1","--bf16"]
parser = options.get_training_parser(); args = options.parse_args_and_arch(parser, input_args=flags)
task = tasks.setup_task(args); model = task.build_model(args); loss = task.build_loss(args)
trainer = Trainer(args, task, model, loss); trainer._total_train_steps = args.max_update
task.load_dataset("train"); ds = task.dataset("train")
batches = [utils.move_to_cuda(ds.collater([ds[k*32+i] for i in range(32)])) for k in range(2)]
for i in range(3): trainer.train_step([batches[i%2]])
torch.cuda.synchronize()
import cProfile, pstats, time
pr = cProfile.Profile(); t0 = time.time(); pr.enable()
for i in range(4): trainer.train_step([batches[i%2]])
torch.cuda.synchronize(); pr.disable()
print("cProfile: %.1f ms/step wall" % ((time.time()-t0)*250))
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(2): trainer.train_step([batches[i%2]])
    torch.cuda.synchronize()
rows=[]
for e in prof.key_averages():
    t=getattr(e,"device_time_total",0)
    if t>0: rows.append((t/2,e.count/2,e.key))
rows.sort(reverse=True); tot=sum(r[0] for r in rows)
print("total device %.1f ms over %d kernels"%(tot/1e3,sum(r[1] for r in rows)))
for t,c,k in rows[:28]: print("%9.1f us %6.1f %5.1f%% %s"%(t,c,100*t/tot,k[:120]))
