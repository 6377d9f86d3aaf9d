"""Paged KV cache for decoder self-attention (north_star (c)); host-side page bookkeeping.

Pool layout (device, `dtype`): [n_pages][dec_layers][heads][k|v][page_tokens][head_dim] (opaque to
the host: include/mdc_b200.h) -- one page holds `page_tokens` consecutive positions of ONE image
for ALL layers, so a single page table serves every layer.  The page table is int32 [B, pages_per_seq] of physical page ids.  Pages are handed
out from a free list; `interleave=True` deals them round-robin across sequences so that a
sequence's pages are deliberately NOT contiguous (exercises the indirection).
"""
import torch


class PageAllocator:
    """Pure host logic (unit-tested on CPU)."""

    def __init__(self, n_pages):
        self.n_pages = int(n_pages)
        self.free = list(range(self.n_pages - 1, -1, -1))
        self.owned = {}

    def alloc(self, seq_id, n):
        if n > len(self.free):
            raise MemoryError(f"KV page pool exhausted: want {n}, free {len(self.free)}")
        pages = [self.free.pop() for _ in range(n)]
        self.owned.setdefault(seq_id, []).extend(pages)
        return pages

    def release(self, seq_id):
        pages = self.owned.pop(seq_id, [])
        self.free.extend(reversed(pages))
        return len(pages)

    @property
    def n_free(self):
        return len(self.free)


def pages_for(tokens, page_tokens):
    return (int(tokens) + page_tokens - 1) // page_tokens


def build_page_table(allocator, batch, max_tokens, page_tokens, interleave=True):
    """Returns a python list-of-lists [B][pages_per_seq] of physical page ids."""
    pps = pages_for(max_tokens, page_tokens)
    table = [[-1] * pps for _ in range(batch)]
    if interleave:
        for j in range(pps):
            for b in range(batch):
                table[b][j] = allocator.alloc(b, 1)[0]
    else:
        for b in range(batch):
            table[b] = allocator.alloc(b, pps)
    return table


class PagedKVCache:
    def __init__(self, batch, max_tokens, layers, dim, page_tokens, dtype, device, interleave=True, spare_pages=0):
        self.page_tokens = page_tokens
        self.pages_per_seq = pages_for(max_tokens, page_tokens)
        n_pages = batch * self.pages_per_seq + spare_pages
        self.allocator = PageAllocator(n_pages)
        self.pool = torch.zeros((n_pages, layers, 2, page_tokens, dim), dtype=dtype, device=device)
        table = build_page_table(self.allocator, batch, max_tokens, page_tokens, interleave)
        self.page_table = torch.tensor(table, dtype=torch.int32, device=device)

    def nbytes(self):
        return self.pool.numel() * self.pool.element_size()
