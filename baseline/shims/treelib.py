"""Minimal stand-in for treelib 1.6.1 (absent from this image), covering exactly the API the
reference's identification code uses (library/identify*.py): Tree.create_node / get_node /
all_nodes / leaves / parent / children / siblings / is_ancestor / paths_to_leaves, Node with
identifier / tag / data, hashable and orderable.  Orderings follow treelib's observable behaviour:
all_nodes() in creation order, children()/siblings()/leaves() in creation order.
Written for the parity harness (baseline/run_pipeline.py); not part of the product.
"""


class Node:
    def __init__(self, tag=None, identifier=None, data=None):
        self.identifier = identifier
        self.tag = identifier if tag is None else tag
        self.data = data
        self._parent = None
        self._children = []

    def __lt__(self, other):
        return self.tag < other.tag

    def __repr__(self):
        return "Node(tag=%s, identifier=%s, data=%s)" % (self.tag, self.identifier, self.data)


class Tree:
    def __init__(self):
        self._nodes = {}
        self.root = None

    def create_node(self, tag=None, identifier=None, parent=None, data=None):
        node = Node(tag=tag, identifier=identifier, data=data)
        if identifier in self._nodes:
            raise ValueError("duplicate node %r" % (identifier,))
        if parent is None:
            if self.root is not None:
                raise ValueError("a tree takes one root only")
            self.root = identifier
        else:
            pid = parent.identifier if isinstance(parent, Node) else parent
            if pid not in self._nodes:
                raise KeyError("parent %r is not in the tree" % (pid,))
            node._parent = pid
            self._nodes[pid]._children.append(identifier)
        self._nodes[identifier] = node
        return node

    def get_node(self, nid):
        return self._nodes.get(nid)

    def __getitem__(self, nid):
        return self._nodes[nid]

    def __contains__(self, nid):
        return nid in self._nodes

    def all_nodes(self):
        return list(self._nodes.values())

    def leaves(self):
        return [n for n in self._nodes.values() if not n._children]

    def parent(self, nid):
        p = self._nodes[nid]._parent
        return None if p is None else self._nodes[p]

    def children(self, nid):
        return [self._nodes[c] for c in self._nodes[nid]._children]

    def siblings(self, nid):
        p = self._nodes[nid]._parent
        if p is None:
            return []
        return [self._nodes[c] for c in self._nodes[p]._children if c != nid]

    def is_ancestor(self, ancestor, grandchild):
        p = self._nodes[grandchild]._parent
        while p is not None:
            if p == ancestor:
                return True
            p = self._nodes[p]._parent
        return False

    def depth(self, node=None):
        if node is None:
            return max((self.depth(n) for n in self._nodes.values()), default=0)
        nid = node.identifier if isinstance(node, Node) else node
        d = 0
        while self._nodes[nid]._parent is not None:
            nid = self._nodes[nid]._parent
            d += 1
        return d

    def paths_to_leaves(self):
        res = []
        for leaf in self.leaves():
            path = [leaf.identifier]
            p = leaf._parent
            while p is not None:
                path.append(p)
                p = self._nodes[p]._parent
            res.append(path[::-1])
        return res
