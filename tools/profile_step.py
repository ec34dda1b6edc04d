"""Three eager (uncaptured) training steps of BASELINE config A for ncu launch lists / full captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pytorch_distributed_nlp_b200 as b2

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
torch.cuda.set_device(0)
cfg = b2.chinese_bert_wwm_ext_config(num_labels=6)
b2.set_seed(123)
model = b2.BertForSequenceClassification(cfg).cuda().train()
args = b2.Args()
opt = b2.build_optimizer(model, args)
step = b2.FusedTrainStep(model, opt, 32, 128, use_graph=False)
for i in range(steps):
    step(b2.synthetic_batch(cfg, 32, 128, 1000 + i))
print("loss", step.loss_to_host())
