"""AggregationInfo: how an aggregate signature was built -- a map
(message_hash, public_key) -> exponent plus the key lists in sorted order
(behaviour of bls_py/aggregation_info.py:7-167).

Host-side bookkeeping only (dict / sort / SHA-256); the reference re-runs two field
inversions for every PublicKey comparison (keys.py:57-64), here keys compare by their cached
serialised bytes."""
from .util import GROUP_ORDER, hash256, hash_pks


class AggregationInfo:
    def __init__(self, tree, message_hashes, public_keys):
        self.tree = tree
        self.message_hashes = message_hashes
        self.public_keys = public_keys

    def empty(self):
        return not self.tree

    # ordering: by the (message hash, public key, exponent) triples, shorter prefix first
    def _triples(self):
        return [(mh, pk, self.tree[(mh, pk)]) for mh, pk in zip(self.message_hashes, self.public_keys)]

    def __lt__(self, other):
        mine, theirs = self._triples(), other._triples()
        for a, b in zip(mine, theirs):
            if a < b:
                return True
            if b < a:
                return False
        return len(mine) < len(theirs)

    def __eq__(self, other):
        return not self < other and not other < self

    def __str__(self):
        return "".join("(%s,%s):\n%s\n" % (mh.hex(), pk.serialize().hex(), hex(e))
                       for (mh, pk), e in self.tree.items())

    def copy(self):
        return AggregationInfo(dict(self.tree), list(self.message_hashes), list(self.public_keys))

    __deepcopy__ = lambda self, memo: self.copy()

    @staticmethod
    def from_msg_hash(public_key, message_hash):
        return AggregationInfo({(message_hash, public_key): 1}, [message_hash], [public_key])

    @staticmethod
    def from_msg(pk, message):
        return AggregationInfo.from_msg_hash(pk, hash256(message))

    @staticmethod
    def _from_tree(tree):
        keys = sorted(tree.keys())
        return AggregationInfo(tree, [k[0] for k in keys], [k[1] for k in keys])

    @staticmethod
    def simple_merge_infos(aggregation_infos):
        """disjoint infos: union of the trees, exponents untouched"""
        tree = {}
        for info in aggregation_infos:
            tree.update(info.tree)
        return AggregationInfo._from_tree(tree)

    @staticmethod
    def secure_merge_infos(colliding_infos):
        """infos sharing messages: info i is raised to T_i = hash_pks(...)[i]"""
        colliding_infos.sort()
        keys = sorted(k for info in colliding_infos for k in info.tree)
        ts = hash_pks(len(colliding_infos), [pk for _, pk in keys])
        tree = {}
        for t, info in zip(ts, colliding_infos):
            for key, exponent in info.tree.items():
                tree[key] = (tree.get(key, 0) + exponent * t) % GROUP_ORDER
        return AggregationInfo._from_tree(tree)

    @staticmethod
    def merge_infos(aggregation_infos):
        seen, colliding = set(), set()
        for info in aggregation_infos:
            local = set(k[0] for k in info.tree)
            colliding |= seen & local
            seen |= local
        if not colliding:
            return AggregationInfo.simple_merge_infos(aggregation_infos)
        hit = [i for i in aggregation_infos if any(k[0] in colliding for k in i.tree)]
        rest = [i for i in aggregation_infos if not any(k[0] in colliding for k in i.tree)]
        rest.append(AggregationInfo.secure_merge_infos(hit))
        return AggregationInfo.simple_merge_infos(rest)
