"""Literal CPU restatement of ``evaluate()`` (train_gnn.py:290-367): per-user python loop,
``torch.mm`` + ``torch.topk`` against the test-split candidate pool, Recall@K and sklearn NDCG@K.
TEST INFRASTRUCTURE (see oracle/__init__.py)."""
from collections import defaultdict

import numpy as np
import torch
from sklearn.metrics import ndcg_score


def evaluate(test_edges, user_emb, post_emb, num_users, K=10):
    user_test_posts = defaultdict(list)
    candidate_posts = set()
    for i in range(test_edges.shape[1]):
        u = test_edges[0, i].item()
        p_local = test_edges[1, i].item() - num_users        # train_gnn.py:316
        user_test_posts[u].append(p_local)
        candidate_posts.add(p_local)
    candidate_posts = torch.tensor(sorted(candidate_posts))
    recall_list, ndcg_list = [], []
    for user_id in user_test_posts:
        if user_id >= num_users:
            continue
        true_posts = user_test_posts[user_id]
        scores = torch.mm(user_emb[user_id].unsqueeze(0), post_emb[candidate_posts].T).squeeze(0)
        topk_idx = torch.topk(scores, min(K, len(scores)))[1]
        topk_posts = candidate_posts[topk_idx].tolist()
        hits = len(set(topk_posts) & set(true_posts))
        recall_list.append(hits / len(true_posts))
        relevance = torch.zeros(len(candidate_posts))
        for p in true_posts:
            idx = (candidate_posts == p).nonzero(as_tuple=True)[0]
            if len(idx) > 0:
                relevance[idx] = 1.0
        if relevance.sum() > 0:
            ndcg_list.append(ndcg_score(relevance.numpy().reshape(1, -1), scores.numpy().reshape(1, -1), k=K))
    return float(np.mean(recall_list)), float(np.mean(ndcg_list))
