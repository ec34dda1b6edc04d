"""ORACLE — test infrastructure only (see bert_ref.py header).  CPU fp32 restatement of the reference step:
bert_ref (HF BertForSequenceClassification math), adamw_ref (HF AdamW), ddp_ref (DDP gradient mean),
cpu_step (the single-gpu-cls.py loop body timed on host cores for bench.py's cpu_baseline / --impl reference)."""
