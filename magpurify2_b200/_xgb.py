"""Evaluation of an XGBoost `gbtree` / `binary:logistic` model saved as JSON (Booster.save_model('*.json')),
without the xgboost package: the reference classifies contigs with a 16-tree / 6-feature model
(`magpurify2/models/fast_model.json`, loaded at magpurify2/core.py:541-546 and applied with
`model.predict(xgb.DMatrix(features))`).  SURVEY.md §8(f) F3.

Semantics restated from xgboost's CPU predictor (float32 throughout, as DMatrix stores float32):
  * a node is a leaf when left_children[n] == -1; its value is split_conditions[n];
  * otherwise x = features[split_indices[n]]; a missing value (NaN) follows default_left[n];
    else the sample goes left when x < split_conditions[n];
  * margin = logit(base_score) + sum of the leaf values in tree order; probability = 1 / (1 + exp(-margin)).
The model is tiny (tens of trees, a few nodes each) and runs once per contig on six numbers: host NumPy."""
import json

import numpy as np


class XgbJsonModel:
    def __init__(self, path):
        with open(path) as fin:
            doc = json.load(fin)
        learner = doc["learner"]
        booster = learner["gradient_booster"]
        if booster.get("name") != "gbtree":
            raise ValueError(f"{path}: only gbtree boosters are supported (found {booster.get('name')!r})")
        objective = learner["objective"]["name"]
        if objective != "binary:logistic":
            raise ValueError(f"{path}: only binary:logistic models are supported (found {objective!r})")
        params = learner["learner_model_param"]
        self.num_feature = int(params["num_feature"])
        self.base_score = np.float32(float(params.get("base_score", "0.5")))
        self.trees = []
        for t in booster["model"]["trees"]:
            self.trees.append({
                "left": np.asarray(t["left_children"], dtype=np.int64),
                "right": np.asarray(t["right_children"], dtype=np.int64),
                "feature": np.asarray(t["split_indices"], dtype=np.int64),
                "cond": np.asarray(t["split_conditions"], dtype=np.float32),
                "default_left": np.asarray(t["default_left"], dtype=bool),
            })

    def margin(self, features):
        x = np.ascontiguousarray(features, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.num_feature:
            raise ValueError(f"expected a [n, {self.num_feature}] feature matrix, got {x.shape}")
        n = x.shape[0]
        rows = np.arange(n)
        with np.errstate(divide="ignore"):
            out = np.full(n, -np.log(np.float32(1.0) / self.base_score - np.float32(1.0)), dtype=np.float32)
        for t in self.trees:
            node = np.zeros(n, dtype=np.int64)
            for _ in range(len(t["left"])):  # a path is shorter than the number of nodes
                inner = t["left"][node] != -1
                if not inner.any():
                    break
                v = x[rows, t["feature"][node]]
                left = np.where(np.isnan(v), t["default_left"][node], v < t["cond"][node])
                node = np.where(inner, np.where(left, t["left"][node], t["right"][node]), node)
            out = (out + t["cond"][node]).astype(np.float32)
        return out

    def predict(self, features):
        m = self.margin(features)
        with np.errstate(over="ignore"):
            return (np.float32(1.0) / (np.float32(1.0) + np.exp(-m, dtype=np.float32))).astype(np.float32)
