"""Subtitle scripts either side of the matcher: ASS and SRT parsing / writing (the data formats of
SURVEY.md section 8f-4).  Behavioural mirror of the reference's subs.py:93-274 on top of ScriptEvent; the
reference's unit tests for it (tests/subtitles.py:22-155) are ported in tests/test_script.py.
"""
import collections
import io
import os
import re

from .common import SushiError, format_time, py2_round
from .events import ScriptEvent


def parse_ass_time(text):
    h, m, s = text.split(':')
    return float(h) * 3600 + float(m) * 60 + float(s)


def format_srt_time(seconds):
    ms = py2_round(seconds * 1000)
    return '{0:02d}:{1:02d}:{2:02d},{3:03d}'.format(int(ms // 3600000), int((ms // 60000) % 60),
                                                    int((ms // 1000) % 60), int(ms % 1000))


class SrtEvent(ScriptEvent):
    __slots__ = ('style',)
    # index, "h:m:s,ms --> h:m:s,ms", then text up to the next such header or the end (subs.py:97-106)
    _STAMP = r'\d{1,2}:\d{1,2}:\d{1,2},\d+'
    EVENT_REGEX = re.compile(r'(\d+?)\s+?(' + _STAMP + r')\s-->\s(' + _STAMP + r').(.+?)'
                             r'(?=(?:\d+?\s+?' + _STAMP + r'\s-->\s' + _STAMP + r')|$)', flags=re.DOTALL)

    def __init__(self, source_index, start, end, text):
        super(SrtEvent, self).__init__(source_index, start, end, text, is_comment=False)
        self.style = None

    @classmethod
    def from_match(cls, match):
        return cls(int(match.group(1)), cls.parse_time(match.group(2)), cls.parse_time(match.group(3)),
                   match.group(4).strip())

    @classmethod
    def from_string(cls, text):
        return cls.from_match(cls.EVENT_REGEX.match(text))

    @staticmethod
    def parse_time(text):
        return parse_ass_time(text.replace(',', '.'))

    def __str__(self):
        return '{0}\n{1} --> {2}\n{3}'.format(self.source_index, format_srt_time(self.start),
                                              format_srt_time(self.end), self.text)


class AssEvent(ScriptEvent):
    __slots__ = ('kind', 'layer', 'style', 'name', 'margin_left', 'margin_right', 'margin_vertical', 'effect')

    def __init__(self, line, position=0):
        kind, _, rest = line.partition(':')
        f = [x.strip() for x in rest.split(',', 9)]          # 9 splits: the text keeps its commas
        super(AssEvent, self).__init__(position, parse_ass_time(f[1]), parse_ass_time(f[2]), f[9],
                                       is_comment=kind.lower() == 'comment')
        self.kind = kind
        self.layer, self.style, self.name = f[0], f[3], f[4]
        self.margin_left, self.margin_right, self.margin_vertical, self.effect = f[5], f[6], f[7], f[8]

    def __str__(self):
        return '{0}: {1},{2},{3},{4},{5},{6},{7},{8},{9},{10}'.format(
            self.kind, self.layer, format_time(self.start), format_time(self.end), self.style, self.name,
            self.margin_left, self.margin_right, self.margin_vertical, self.effect, self.text)


class _Script(object):
    def __init__(self, events):
        self.events = events

    def sort_by_time(self):
        self.events.sort(key=lambda e: e.start)


class SrtScript(_Script):
    @classmethod
    def from_file(cls, path):
        try:
            with io.open(path, encoding='utf-8-sig') as f:
                text = f.read()
        except IOError:
            raise SushiError('Script {0} not found'.format(path))
        return cls([SrtEvent.from_match(m) for m in SrtEvent.EVENT_REGEX.finditer(text)])

    def save_to_file(self, path):
        with io.open(path, 'w', encoding='utf-8', newline='') as f:
            f.write('\n\n'.join(str(e) for e in self.events))


class AssScript(_Script):
    STYLES_FORMAT = ('Format: Name, Fontname, Fontsize, PrimaryColour, SecondaryColour, OutlineColour, BackColour, '
                     'Bold, Italic, Underline, StrikeOut, ScaleX, ScaleY, Spacing, Angle, BorderStyle, Outline, '
                     'Shadow, Alignment, MarginL, MarginR, MarginV, Encoding')
    EVENTS_FORMAT = 'Format: Layer, Start, End, Style, Name, MarginL, MarginR, MarginV, Effect, Text'

    def __init__(self, script_info, styles, events, other):
        super(AssScript, self).__init__(events)
        self.script_info, self.styles, self.other = script_info, styles, other

    @classmethod
    def from_file(cls, path):
        info, styles, events = [], [], []
        other = collections.OrderedDict()
        known = {'[script info]': info.append, '[v4+ styles]': styles.append,
                 '[events]': lambda line: events.append(AssEvent(line, position=len(events) + 1))}
        sink = None
        try:
            with io.open(path, encoding='utf-8-sig') as f:
                for number, raw in enumerate(f):
                    line = raw.strip()
                    if not line:
                        continue
                    low = line.lower()
                    if low in known:
                        sink = known[low]
                    elif re.match(r'\[.+?\]', low):
                        if line in other:
                            raise SushiError('Duplicate section detected, invalid script?')
                        other[line] = []
                        sink = other[line].append
                    elif sink is None:
                        raise SushiError("That's some invalid ASS script")
                    elif sink in known.values() and line.startswith('Format:'):
                        continue                                   # format lines are regenerated on save
                    else:
                        try:
                            sink(line)
                        except Exception as e:
                            raise SushiError("That's some invalid ASS script: {0} [line {1}]".format(e, number))
        except IOError:
            raise SushiError('Script {0} not found'.format(path))
        return cls(info, styles, events, other)

    def save_to_file(self, path):
        out = []
        if self.script_info:
            out += ['[Script Info]'] + list(self.script_info) + ['']
        if self.styles:
            out += ['[V4+ Styles]', self.STYLES_FORMAT] + list(self.styles) + ['']
        if self.events:
            out += ['[Events]', self.EVENTS_FORMAT] + [str(e) for e in sorted(self.events, key=lambda e: e.source_index)]
        for name, lines in (self.other or {}).items():
            out += ['', name] + list(lines)
        with io.open(path, 'w', encoding='utf-8-sig', newline='') as f:
            f.write(os.linesep.join(out))


def load_script(path):
    ext = os.path.splitext(path)[1].lower()
    if ext == '.ass':
        return AssScript.from_file(path)
    if ext == '.srt':
        return SrtScript.from_file(path)
    raise SushiError('Unknown script type')
