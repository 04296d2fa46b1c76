class MPSBatch: pass
