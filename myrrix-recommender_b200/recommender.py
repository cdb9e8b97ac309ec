"""Host mirror of the reference's recommend calls over the resident model.

Same method names and argument meaning as net.myrrix.online.ServerRecommender
(online/src/net/myrrix/online/ServerRecommender.java): recommend (:355-358), recommendToMany
(:366-441), recommendToAnonymous (:511-560).  IDs are the caller's long IDs; the dense-index remap
is the shim's.  Scoring and selection run on the GPU (csrc/topn.cuh through
include/myrrix_als.h); nothing here computes a score.
"""
import numpy as np


class NoSuchUserException(KeyError):
    """org.apache.mahout.cf.taste.common.NoSuchUserException"""


class NoSuchItemException(KeyError):
    """org.apache.mahout.cf.taste.common.NoSuchItemException"""


class NotReadyException(RuntimeError):
    """net.myrrix.common.NotReadyException"""


class Recommender:
    """A built model resident on the device: `als` is the NativeALS handle that holds X, Y and R;
    user_ids / item_ids give the long ID of every dense row (first-appearance order of the
    ingest library)."""

    def __init__(self, als, user_ids, item_ids, foldin=None):
        self.als = als
        self.user_ids = np.asarray(user_ids, dtype=np.int64)
        self.item_ids = np.asarray(item_ids, dtype=np.int64)
        self._user_index = {int(u): i for i, u in enumerate(self.user_ids)}
        self._item_index = {int(v): i for i, v in enumerate(self.item_ids)}
        self.foldin = foldin  # myrrix_recommender_b200.foldin.FoldIn (recommendToAnonymous)

    def _out(self, items, values):
        return [(int(self.item_ids[i]), float(v)) for i, v in zip(items, values)]

    def recommend(self, userID, howMany, considerKnownItems=False, excludeItemIDs=()):
        return self.recommendToMany([userID], howMany, considerKnownItems, excludeItemIDs)

    def recommendToMany(self, userIDs, howMany, considerKnownItems=False, excludeItemIDs=()):
        if howMany <= 0:
            raise ValueError("howMany must be positive")  # ServerRecommender.java:371
        rows = [self._user_index[int(u)] for u in userIDs if int(u) in self._user_index]
        if not rows:
            raise NoSuchUserException(str(list(userIDs)))  # :391-393
        ex = [self._item_index[int(i)] for i in excludeItemIDs if int(i) in self._item_index]
        return self._out(*self.als.recommend(rows, howMany, considerKnownItems, ex))

    def setPreference(self, userID, itemID, value=1.0):
        """The model part of ServerRecommender.setPreference (:735-760) for a known user and item: the
        write is folded into the resident rows (updateFeatures); new IDs wait for the next build."""
        if int(userID) not in self._user_index or int(itemID) not in self._item_index:
            return False
        self.als.fold_in([self._user_index[int(userID)]], [self._item_index[int(itemID)]], [value])
        return True

    def recommendToAnonymous(self, itemIDs, values=None, howMany=10):
        """buildAnonymousUserFeatures (ServerRecommender.java:561-608, host fold-in library), then
        the same scoring with the anonymous user's items filtered (:540-552)."""
        if self.foldin is None:
            raise NotReadyException("no fold-in state: build it from the model's Gramians first")
        known = [self._item_index[int(i)] for i in itemIDs if int(i) in self._item_index]
        if not known:
            raise NoSuchItemException(str(list(itemIDs)))
        vals = np.ones(len(itemIDs), np.float32) if values is None else np.asarray(values, np.float32)
        vals = np.asarray([v for i, v in zip(itemIDs, vals) if int(i) in self._item_index], np.float32)
        rows = self.als.get_rows("y", known)
        feats = self.foldin.anonymous_user(rows, vals)
        return self._out(*self.als.top_n("y", feats, howMany, exclude=known))
