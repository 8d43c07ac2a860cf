"""Minimal subtitle-event object with the shift/diff/link protocol that the shift solver and the
grouping heuristics use (the fields of the reference's subs.ScriptEventBase, subs.py:14-83, that this
path touches: start, end, shift, diff, link chain).  Script parsing/writing is out of scope."""


class ScriptEvent(object):
    __slots__ = ('source_index', 'start', 'end', 'text', 'is_comment', '_shift', '_diff', '_link',
                 '_start_shift', '_end_shift')

    def __init__(self, source_index, start, end, text='', is_comment=False):
        self.source_index = source_index
        self.start = start
        self.end = end
        self.text = text
        self.is_comment = is_comment
        self._shift = 0
        self._diff = 1
        self._link = None
        self._start_shift = 0
        self._end_shift = 0

    # values resolve through the link chain (subs.py:27-33)
    @property
    def linked(self):
        return self._link is not None

    @property
    def shift(self):
        return self._link.shift if self._link is not None else self._shift

    @property
    def diff(self):
        return self._link.diff if self._link is not None else self._diff

    @property
    def duration(self):
        return self.end - self.start

    @property
    def shifted_start(self):
        return self.start + self.shift + self._start_shift

    @property
    def shifted_end(self):
        return self.end + self.shift + self._end_shift

    def set_shift(self, shift, audio_diff):
        assert self._link is None, 'Cannot set shift of a linked event'
        self._shift = shift
        self._diff = audio_diff

    def adjust_shift(self, value):
        assert self._link is None, 'Cannot adjust time of linked events'
        self._shift += value

    def adjust_additional_shifts(self, start_shift, end_shift):
        assert self._link is None, 'Cannot apply additional shifts to a linked event'
        self._start_shift += start_shift
        self._end_shift += end_shift

    def get_link_chain_end(self):
        node = self
        while node._link is not None:
            node = node._link
        return node

    def link_event(self, other):
        assert other.get_link_chain_end() is not self, 'Circular link detected'
        self._link = other

    def resolve_link(self):
        assert self._link is not None, 'Cannot resolve unlinked events'
        self._shift, self._diff = self._link.shift, self._link.diff
        self._link = None

    def apply_shift(self):
        self.start, self.end = self.shifted_start, self.shifted_end

    def __repr__(self):
        return 'ScriptEvent(#{0} {1:.2f}-{2:.2f} shift={3})'.format(self.source_index, self.start, self.end, self.shift)
