"""Literal Python transcription of the DBoW2 transform used by Frame::ComputeBoW
(Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1139-1203,1230-1271; BowVector.cpp:34-46,62-84;
FeatureVector.cpp:31-45), TF_IDF + L1: pins the oracle."""
import numpy as np


def dist(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def transform(voc, features, levelsup=4):
    ptr, idx, desc, word_id, weight, L = voc["child_ptr"], voc["child_idx"], voc["node_desc"], voc["word_id"], voc["weight"], voc["L"]
    v, fv = {}, {}
    words, nodes = [], []
    for i_feature, f in enumerate(features):
        nid_level = L - levelsup
        nid = 0
        final_id = 0
        level = 0
        while True:
            level += 1
            kids = idx[ptr[final_id]:ptr[final_id + 1]]
            final_id = int(kids[0])
            best = float(dist(f, desc[final_id]))
            for c in kids[1:]:
                d = float(dist(f, desc[int(c)]))
                if d < best:
                    best, final_id = d, int(c)
            if level == nid_level:
                nid = final_id
            if ptr[final_id] == ptr[final_id + 1]:
                break
        w = float(weight[final_id])
        words.append(int(word_id[final_id]))
        nodes.append(nid)
        if w > 0:
            v[int(word_id[final_id])] = v.get(int(word_id[final_id]), 0.0) + w
            fv.setdefault(nid, []).append(i_feature)
    norm = 0.0
    for k in sorted(v):
        norm += abs(v[k])
    if norm > 0.0:
        for k in v:
            v[k] /= norm
    return words, nodes, sorted(v.items()), sorted(fv.items())
