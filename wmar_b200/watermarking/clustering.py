"""CLUSTERING greenlist split (wmar/watermarking/gentime_watermark.py:175-216), host side.

The reference builds this split once per watermarker (fixed seeding only): t-SNE of the alive codebook vectors to 2-D,
KMeans into 100 clusters, clusters coloured like a chess board (rows of ten centres sorted by y, inside a row by x,
colours alternating), green = alive ids of the clusters coloured 1 plus the even dead ids.  It is a one-off host
computation on the codebook (seconds to minutes in scikit-learn), not part of the per-token path, so it stays on the
host here too and its result becomes the single row of the device bitmask table.  scikit-learn is the same third-party
dependency the reference imports (TSNE / KMeans with random_state 42): identical library, identical calls.
"""
import numpy as np


def clustering_greenlist_ids(embedding, alive_ids, dead_ids, n_clusters=100):
    """embedding [vocab, dim] (torch tensor or ndarray), alive_ids / dead_ids sequences of ints -> list of green ids in
    the reference's order (alive ids in file order, then the even dead ids)."""
    from sklearn.cluster import KMeans
    from sklearn.manifold import TSNE

    emb = embedding.detach().cpu().numpy() if hasattr(embedding, "detach") else np.asarray(embedding)
    alive = [int(i) for i in (alive_ids.tolist() if hasattr(alive_ids, "tolist") else alive_ids)]
    dead = [int(i) for i in (dead_ids.tolist() if hasattr(dead_ids, "tolist") else dead_ids)]
    flat = emb[alive].reshape(len(alive), -1)
    pts = TSNE(n_components=2, random_state=42).fit_transform(flat)
    kmeans = KMeans(n_clusters=n_clusters, random_state=42)
    kmeans.fit(pts)
    centers = kmeans.cluster_centers_
    labels = np.arange(len(centers))
    ysort = np.argsort(centers[:, 1])                     # rows of the board: by y
    centers, labels = centers[ysort], labels[ysort]
    centers = centers.reshape(-1, 10, 2)
    labels = labels.reshape(-1, 10)
    colour = {}
    curr = 0
    for i in range(centers.shape[0]):
        curr = 1 - curr                                   # every row starts with the colour the previous row ended on
        xsort = np.argsort(centers[i, :, 0])
        for lab in labels[i][xsort]:
            colour[int(lab)] = curr
            curr = 1 - curr
    green = [idd for i, idd in enumerate(alive) if colour[int(kmeans.labels_[i])] == 1]
    green += [idd for idd in dead if idd % 2 == 0]
    return green


def ids_to_bitmask_row(ids, vocab_size):
    """int32[ceil(V / 32)] little-endian bit row (bit v set <=> v is green), the layout of the device table."""
    bits = np.zeros(((vocab_size + 31) // 32) * 32, dtype=np.uint8)
    bits[np.asarray(ids, dtype=np.int64)] = 1
    return np.packbits(bits, bitorder="little").view(np.int32).copy()
