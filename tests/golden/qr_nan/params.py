class Quadrupolar:
    data_set = './efg.csv'
    analyte = '127I'
