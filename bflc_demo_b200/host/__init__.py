"""Host-side (CPU) client runtime: the reference's L3/L4 layers (python-sdk/main.py) on top of
the C++ ledger -- in-process simulator, gloo-replicated multi-process path and an RPC ledger
service with a process launcher."""
