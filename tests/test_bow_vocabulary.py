"""next row N2: DBoW2 vocabulary-tree descent (TemplatedVocabulary::transform, Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h
:1119-1259) — oracle, text-file loader and host-side BowVector/FeatureVector bookkeeping on CPU; CUDA descent on GPU."""
import numpy as np
import pytest


def _ref_transform(tree, desc, levelsup):
    """independent pure-Python restatement of the descent, to cross-check the C oracle"""
    cs, ci, nd, L = tree['child_start'], tree['child_ids'], tree['desc'], tree['L']
    out = []
    for d in desc:
        cur, level, nid = 0, 0, 0
        while True:
            ch = ci[cs[cur]:cs[cur + 1]]
            dist = [int(np.unpackbits(d ^ nd[c]).sum()) for c in ch]
            cur = int(ch[int(np.argmin(dist))])            # argmin returns the first minimum = strict '<' scan
            level += 1
            if level == L - levelsup:
                nid = cur
            if cs[cur + 1] == cs[cur]:
                break
        out.append((int(tree['word'][cur]), nid, float(tree['weight'][cur])))
    return out


def test_oracle_descent_and_text_loader(pkg, oracle, synth, tmp_path):
    for k, L, ragged in ((10, 3, False), (6, 4, True)):
        tree, lines = synth.synthetic_vocabulary(k, L, seed=5 + k, ragged=ragged)
        desc = synth.random_descriptors(9, 300)
        desc[:150] = synth.flip_bits(tree['desc'][np.random.default_rng(1).integers(1, len(tree['word']), 150)], 33, [20] * 150)
        wid, nid, w = oracle.bow_transform(tree, desc, levelsup=L - 1)
        ref = _ref_transform(tree, desc[:60], L - 1)
        assert [(int(a), int(b), float(c)) for a, b, c in zip(wid[:60], nid[:60], w[:60])] == ref
        path = tmp_path / ('voc_%d.txt' % k)
        path.write_text('\n'.join(lines) + '\n')
        parsed = pkg.ORBVocabulary.parse_text(str(path))
        for key in ('child_start', 'child_ids', 'word'):
            assert np.array_equal(parsed[key], tree[key]), key
        assert np.array_equal(parsed['desc'][1:], tree['desc'][1:])      # the root has no line in the file (and no descriptor)
        assert np.array_equal(parsed['weight'], tree['weight']) and (parsed['k'], parsed['L']) == (k, L)
        bow, fv = pkg.ORBVocabulary.accumulate(wid, nid, w, 0, 0)
        assert abs(sum(bow.values()) - 1.0) < 1e-12          # L1-normalised TF-IDF
        assert sorted(i for v in fv.values() for i in v) == [i for i in range(len(desc)) if w[i] > 0]
        assert all(v == sorted(v) for v in fv.values())


@pytest.mark.gpu
def test_gpu_descent_matches_oracle(pkg, oracle, synth):
    for k, L, ragged, n in ((10, 4, False, 3000), (7, 5, True, 2000), (20, 2, False, 500), (40, 2, True, 300)):
        tree, _ = synth.synthetic_vocabulary(k, L, seed=100 + k, ragged=ragged)
        desc = synth.random_descriptors(19, n)
        src = np.random.default_rng(2).integers(1, len(tree['word']), n // 2)
        desc[:n // 2] = synth.flip_bits(tree['desc'][src], 44, np.random.default_rng(3).integers(0, 60, n // 2).tolist())
        voc = pkg.ORBVocabulary(tree)
        for levelsup in (0, 1, L - 1, L, L + 2):
            wid, nid, w = voc.descend(desc, levelsup)
            owid, onid, ow = oracle.bow_transform(tree, desc, levelsup)
            assert np.array_equal(wid, owid) and np.array_equal(nid, onid) and np.array_equal(w, ow), (k, L, levelsup)
        bow, fv = voc.transform(desc, 4 if L > 4 else L - 1)
        assert bow and fv
        voc.close()
