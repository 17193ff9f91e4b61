"""TEST INFRASTRUCTURE -- empty stand-in so /root/reference/graph_ltpl/testing_tools/src/objectlist_dummy.py:2 imports."""
PUB = 1


class Context(object):
    def socket(self, *a, **k):
        raise RuntimeError("zmq shim: sockets are not available")
