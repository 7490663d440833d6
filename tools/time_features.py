import sys, torch, numpy as np
sys.path.insert(0, ".")
import bench
from speech_emotion_privacy_trust_b200 import extraction
dev = torch.device("cuda", 0)
lengths = bench.corpus_lengths(5531, 1234)
off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
batch = extraction.RaggedAudio(bench.synth_corpus_device(lengths, 4321, dev), off)
hours = float(lengths.sum())/16000/3600
r = bench.other_features(batch, hours, dev)
print({k: (round(v["ms"], 3), round(v["fp32_frac"], 3)) for k, v in r.items()})
