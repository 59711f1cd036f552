// Per-game LRU cache of network evaluations, keyed by the board hash (rust/kz-selfplay/src/server/generator_alphazero.rs:68,77-79
// keeps an `lru::LruCache<B, ZeroEvaluation>` per game).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "mcts.hpp"

namespace kzb {
namespace selfplay {

// Flat storage: entries live in one array, linked into the recency list and into hash chains by index, and keep their
// policy vector when they are recycled -- a lookup or an insert allocates nothing once the cache is warm.
class LruCache {
public:
    struct Entry {
        uint64_t key = 0;
        int32_t prev = -1, next = -1, chain = -1;
        ValuesPov values;
        std::vector<float> policy;
    };
    explicit LruCache(size_t cap) : cap_(cap) {
        size_t buckets = 16;
        while (buckets < cap * 2) buckets *= 2;
        mask_ = buckets - 1;
        if (cap_) heads_.assign(buckets, Bucket{});
        entries_.reserve(cap_);
    }
    void prefetch(uint64_t key) const {
        if (cap_) __builtin_prefetch(&heads_[size_t(key) & mask_]);
    }
    const Entry* get(uint64_t key) {
        if (cap_ == 0) return nullptr;
        const int32_t i = find(key);
        if (i < 0) return nullptr;
        touch(i);
        return &entries_[size_t(i)];
    }
    // the entry for `key`, made most recent; the caller fills `values` and `policy`.  nullptr when the cache is disabled.
    Entry* put(uint64_t key) {
        if (cap_ == 0) return nullptr;
        int32_t i = find(key);
        if (i >= 0) {
            touch(i);
            return &entries_[size_t(i)];
        }
        if (entries_.size() < cap_) {
            i = int32_t(entries_.size());
            entries_.emplace_back();
        } else {  // recycle the least recently used entry
            i = tail_;
            unlink(i);
            unchain(i);
        }
        Entry& e = entries_[size_t(i)];
        e.key = key;
        Bucket& b = heads_[size_t(key) & mask_];
        e.chain = b.head;
        b.head = i;
        b.tags |= tag_bit(key);
        push_front(i);
        return &e;
    }
    void clear() {
        std::fill(heads_.begin(), heads_.end(), Bucket{});
        entries_.clear();
        head_ = tail_ = -1;
    }

private:
    // a bucket carries a 32-bit Bloom word of the keys chained in it: most misses are answered from the bucket's own
    // cache line without touching an entry
    struct Bucket {
        int32_t head = -1;
        uint32_t tags = 0;
    };
    static uint32_t tag_bit(uint64_t key) { return 1u << ((key >> 40) & 31); }
    int32_t find(uint64_t key) const {
        const Bucket& b = heads_[size_t(key) & mask_];
        if (!(b.tags & tag_bit(key))) return -1;
        for (int32_t i = b.head; i >= 0; i = entries_[size_t(i)].chain)
            if (entries_[size_t(i)].key == key) return i;
        return -1;
    }
    void unchain(int32_t i) {
        Bucket& b = heads_[size_t(entries_[size_t(i)].key) & mask_];
        int32_t* link = &b.head;
        while (*link != i) link = &entries_[size_t(*link)].chain;
        *link = entries_[size_t(i)].chain;
        b.tags = 0;  // rebuild the Bloom word from what is left in the chain
        for (int32_t j = b.head; j >= 0; j = entries_[size_t(j)].chain) b.tags |= tag_bit(entries_[size_t(j)].key);
    }
    void unlink(int32_t i) {
        Entry& e = entries_[size_t(i)];
        if (e.prev >= 0) entries_[size_t(e.prev)].next = e.next; else head_ = e.next;
        if (e.next >= 0) entries_[size_t(e.next)].prev = e.prev; else tail_ = e.prev;
    }
    void push_front(int32_t i) {
        Entry& e = entries_[size_t(i)];
        e.prev = -1;
        e.next = head_;
        if (head_ >= 0) entries_[size_t(head_)].prev = i;
        head_ = i;
        if (tail_ < 0) tail_ = i;
    }
    void touch(int32_t i) {
        if (head_ == i) return;
        unlink(i);
        push_front(i);
    }
    size_t cap_, mask_ = 0;
    int32_t head_ = -1, tail_ = -1;
    std::vector<Bucket> heads_;
    std::vector<Entry> entries_;
};

}  // namespace selfplay
}  // namespace kzb
